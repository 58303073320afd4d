"""fish_tts_b200 -- the B200-native dual-AR decode path of smolGura/fish-tts behind the reference's own API.

Re-exports what ``fish_tts/__init__.py:34-37`` re-exports, so ``from fish_tts_b200 import get_instance`` stands where
``from fish_tts import get_instance`` stood.  Importing the package does not touch the GPU or load libdualar.so; the first
engine construction does (and raises when the library or a CUDA device is missing -- there is no fallback path).
"""

__all__ = ["FishTTS", "VoiceProfile", "get_instance", "reset_instance"]


def __getattr__(name):      # lazy: `import fish_tts_b200` stays cheap (no torch import) for the build / symbol checks
    if name in __all__:
        from . import synthesizer
        return getattr(synthesizer, name)
    raise AttributeError(f"module 'fish_tts_b200' has no attribute {name!r}")
