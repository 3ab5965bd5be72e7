"""The reference's test files run unmodified against the `torch_nf` shim (= torch_nf_b200): every item needs the
GPU.  Nothing is skipped or expected to fail: the mixture-of-Gaussians estimator (`test_MoG`, the last block of
`test_ConditionalDensityEstimator`) is built too."""
import pytest


def pytest_collection_modifyitems(config, items):
    for item in items:
        if "reference_suite" not in str(item.fspath):
            continue
        item.add_marker(pytest.mark.gpu)
