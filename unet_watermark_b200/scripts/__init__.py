"""Batched B200 versions of the reference's other callers of the model seam."""
