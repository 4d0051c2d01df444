import os
import sys

import pytest

# The GPU suite interleaves large fp32 oracle temporaries (GBs per attention) with the product's long-lived packed weights and
# CUDA-graph pools; with the default allocator the long-lived blocks pin the large segments (175 GB reserved for 30 GB
# allocated, measured).  Expandable segments return unused pages.  Must be set before torch initialises CUDA.
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "remote-sensing-vision-language-diffusion-model_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _release_cuda_cache(request):
    """GPU tests build several engines (CUDA graphs with private pools) and run the fp32 oracle beside them: return cached
    blocks to the device between tests so a long session does not fragment its way to an out-of-memory error."""
    yield
    if "gpu" in request.keywords:
        import gc

        import torch

        gc.collect()
        if torch.cuda.is_available():
            torch.cuda.empty_cache()
            if os.environ.get("B200SR_TEST_MEM"):
                print(f"\n[mem] after {request.node.name}: allocated {torch.cuda.memory_allocated() / 2**30:.1f} GiB, "
                      f"reserved {torch.cuda.memory_reserved() / 2**30:.1f} GiB")
