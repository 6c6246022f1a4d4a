"""Opt-in kernels, exercised last (file name) and in a child process (their switches are read once per process)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_cta_pair_kernel_opt_in():
    """The cta_group::2 form of the int8 kernel (H2_BM_PAIR=1, csrc/bitmap_mma.cu: bm_mma_pair_kernel) is opt-in; the switch
    is read once per process, so it is exercised in a child process (tools/pair_check.py): north-star graph, 1e-4 against
    the fp32 CSR path.  Whether it also equals the single-CTA kernel bit for bit is reported, not required: the integer
    accumulation is exact in both, but the two stream-K schedules cut the units at different points, so the fp32 partial
    tiles are summed in different groupings."""
    import os
    import subprocess
    import sys
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, H2_BM_PAIR="1")
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "pair_check.py")], env=env, capture_output=True, text=True,
                       timeout=300)
    assert r.returncode in (0, 4), r.stdout + r.stderr
    import re
    m = re.search(r"pair kernel vs csr: rel err ([0-9.eE+-]+)(.*)", r.stdout)
    assert m, r.stdout + r.stderr
    assert float(m.group(1)) <= 1e-4 and "nan" not in m.group(2), r.stdout
