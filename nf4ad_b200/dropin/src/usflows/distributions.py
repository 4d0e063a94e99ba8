from nf4ad_b200.distributions import Normal  # noqa: F401
