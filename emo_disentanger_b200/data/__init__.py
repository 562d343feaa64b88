from .token_store import Stage2TokenStore, Stage1TokenStore  # noqa: F401
