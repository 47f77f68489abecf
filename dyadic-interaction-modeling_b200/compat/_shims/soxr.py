"""Stand-in for `soxr`: transformers.audio_utils imports it whenever a `librosa` module is importable (here: the librosa stand-in
next to this file); nothing on the listener eval path resamples raw audio."""


def __getattr__(name):
    raise NotImplementedError(f"soxr.{name}: soxr is not available offline")
