"""The word-directory variant of the filtered kernel (whitelists of millions of entries take it by
default: test_gpu_match.py::test_dense_index_3M_sized_whitelist_vs_oracle) forced onto small,
tie-rich and crowded whitelists and the 737K list.  The switch is read once per process, hence the
child process (tests/dir_parity.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_word_directory_variant_vs_oracle(cuda_device):
    env = dict(os.environ, NR_FILTER_DIR="1")
    r = subprocess.run([sys.executable, os.path.join(HERE, "dir_parity.py")], env=env, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0 and "dir_parity ok" in r.stdout, (r.stdout[-1500:], r.stderr[-3000:])
