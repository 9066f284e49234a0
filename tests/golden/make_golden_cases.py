"""Case tables shared by the golden generator (tests/golden/make_golden.py) and the tests that re-create the inputs."""

PRE_CASES = [  # (source h, w, seed, letterbox kwargs)
    (60, 90, 1, dict(new_shape=(96, 128), auto=False)),
    (123, 77, 2, dict(new_shape=(96, 128), auto=False)),
    (200, 150, 3, dict(new_shape=(96, 128), auto=False)),          # down-scaling
    (48, 64, 4, dict(new_shape=(96, 128), auto=False)),            # exact 2x up-scaling
    (96, 128, 5, dict(new_shape=(96, 128), auto=False)),           # no resize at all
    (70, 101, 6, dict(new_shape=160, auto=True, stride=32)),       # minimum-rectangle padding
    (70, 101, 7, dict(new_shape=(96, 128), auto=False, scaleup=False)),
    (50, 120, 8, dict(new_shape=(64, 64), auto=False, scaleFill=True)),
]
