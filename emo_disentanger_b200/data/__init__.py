from .token_store import Stage2TokenStore  # noqa: F401
