from micformer_b200.models.MICFormer_self import *  # noqa: F401,F403
