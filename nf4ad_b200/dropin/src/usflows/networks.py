from nf4ad_b200.nn import ConvNet  # noqa: F401
