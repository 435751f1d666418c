from .notes import PianoNote  # noqa: F401
