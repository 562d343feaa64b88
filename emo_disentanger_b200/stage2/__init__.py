from .music_performer import MusicPerformer  # noqa: F401
