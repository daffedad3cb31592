"""Validation-time twins of the post-processing (reference dataset/utils.py)."""
