from nf4ad_b200.optim import SophiaG  # noqa: F401
