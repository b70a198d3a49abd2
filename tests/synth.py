from sgdnet_b200.synth import *  # noqa: F401,F403  (generators live in the package so bench.py can use them)
