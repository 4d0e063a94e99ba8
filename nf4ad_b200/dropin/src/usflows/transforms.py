from nf4ad_b200.transforms import (  # noqa: F401
    BaseTransform, BlockAffineTransform, HouseholderTransform, InverseTransform, LUTransform,
    MaskedCoupling, ScaleTransform, SequentialAffineTransform,
)
