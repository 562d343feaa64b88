from .plain_transformer import PlainTransformer  # noqa: F401
