from micformer_b200.loss.dice import *  # noqa: F401,F403
