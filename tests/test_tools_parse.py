"""Every script under tools/ and oracle/ at least parses (several only run under torchrun on multi-GPU boxes or under
gpurun, where a typo would cost a GPU call)."""
import ast
import glob
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_python_tools_parse():
    files = sorted(glob.glob(os.path.join(ROOT, "tools", "*.py")) + glob.glob(os.path.join(ROOT, "oracle", "*.py")) +
                   [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")])
    assert len(files) > 10
    for f in files:
        ast.parse(open(f).read(), filename=f)


def test_shell_tools_have_no_syntax_errors():
    import subprocess
    for f in sorted(glob.glob(os.path.join(ROOT, "tools", "*.sh"))):
        r = subprocess.run(["bash", "-n", f], capture_output=True, text=True)
        assert r.returncode == 0, (f, r.stderr)
