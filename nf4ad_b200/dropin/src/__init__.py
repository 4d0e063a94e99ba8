"""Import root `src` of the USFlows package (`from src.usflows.flows import Flow`), served by nf4ad_b200."""
