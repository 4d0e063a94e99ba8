from nf4ad_b200.nn import ConditionalDenseNN, DenseNN  # noqa: F401
