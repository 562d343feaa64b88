from .music_performer import MusicPerformer  # noqa: F401
from .music_gpt2 import MusicGPT2  # noqa: F401
