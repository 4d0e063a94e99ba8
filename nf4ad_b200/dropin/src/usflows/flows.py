from nf4ad_b200.flows import Flow, USFlow  # noqa: F401
