"""Stand-in for `librosa`: imported at module level by the reference's dataset modules, used only when raw audio is loaded."""


def __getattr__(name):
    raise NotImplementedError(f"librosa.{name}: librosa is not available offline; the listener eval path reads pre-extracted features")
