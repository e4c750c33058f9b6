from micformer_b200.models.STN import *  # noqa: F401,F403
