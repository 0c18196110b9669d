"""The reference's public/*.py model-class surface, re-implemented over the B200 engine."""
