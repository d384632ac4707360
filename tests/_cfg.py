"""Fixture ``cfg`` dict -> npf_b200 model: lives in the package (npf_b200/utils/configs.py) so that the examples do not depend on the
test tree; re-exported here for the tests."""
from npf_b200.utils.configs import R_DIM, build_model, loss_for  # noqa: F401
