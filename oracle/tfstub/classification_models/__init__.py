"""Placeholder for the un-vendored `classification_models` pip dependency the reference's
RetinaNet backbone uses (RetinaNet/retinanet_module.py:6).  TEST INFRASTRUCTURE ONLY."""
