from .notes import PianoNote  # noqa: F401
from .module import ShaderPiano  # noqa: F401
