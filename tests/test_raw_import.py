"""CPU: the torch-free front end imports without torch and finds the CUDA runtime (no device calls)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_raw_module_does_not_import_torch():
    code = ("import sys; sys.path.insert(0, %r); import paddle3d_b200.raw as r; "
            "assert 'torch' not in sys.modules; r.cudart(); print(sorted(n for n in dir(r) if not n.startswith('_'))[:3])" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
