from nf4ad_b200.nn import DenseNN  # noqa: F401
