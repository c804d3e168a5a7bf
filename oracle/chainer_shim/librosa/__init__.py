"""Import stub: the reference's utils.py does `import librosa` at module level (utils.py:6) but
the hot path never calls it."""
