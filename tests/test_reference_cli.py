"""The reference's own generate.py, byte for byte (baseline/_ref/generate.py, installed by tools/install_reference.sh;
/root/reference/generate.py in the build container), executed against this package through
cold_compress_b200.run_reference_script: its imports of cache / model / generation_utils / tp / attention_utils /
prompt_compression bind to the B200 implementations, its tokenizer stays its own (reference generate.py:54-204)."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _script():
    for p in (ROOT / "baseline" / "_ref" / "generate.py", Path("/root/reference/generate.py")):
        if p.exists():
            return p
    return None


def _run(args, timeout=600):
    env = dict(os.environ, PYTHONPATH=str(ROOT))
    return subprocess.run([sys.executable, "-m", "cold_compress_b200.run_reference_script", *map(str, args)],
                          capture_output=True, text=True, timeout=timeout, cwd=ROOT, env=env)


def test_generate_py_cli_parses_against_this_package():
    """CPU: the unmodified script imports everything it needs from the drop-in modules and reaches its own argument
    check (it asserts that the checkpoint exists before touching a device)."""
    script = _script()
    if script is None:
        pytest.skip("no reference checkout (run tools/install_reference.sh)")
    r = _run([script, "--checkpoint_path", "/nonexistent/Llama-2-stories110M/model.pth", "--cache_config", "heavy_hitter",
              "--max_cache_length", "64", "--prompt", "hello"])
    assert r.returncode != 0
    assert "compatible with each other" in r.stdout, r.stdout[-500:] + r.stderr[-1500:]   # our cache_compatibility ran
    assert "AssertionError" in r.stderr and "model.pth" in r.stderr, r.stderr[-1500:]     # generate.py:68


@pytest.mark.gpu
@pytest.mark.parametrize("extra", [[], ["--compile"]])
def test_generate_py_runs_unmodified_on_the_gpu(tmp_path, extra):
    script = _script()
    if script is None:
        pytest.skip("no reference checkout (run tools/install_reference.sh)")
    w = subprocess.run([sys.executable, str(ROOT / "tools" / "write_random_checkpoint.py"), str(tmp_path), "--model", "stories110M"],
                       capture_output=True, text=True, timeout=600)
    assert w.returncode == 0, w.stderr[-2000:]
    ckpt = Path(w.stdout.strip().splitlines()[-1])
    r = _run([script, "--checkpoint_path", ckpt, "--cache_config", "heavy_hitter", "--max_cache_length", "64",
              "--max_new_tokens", "8", "--prompt", "the cache keeps heavy hitters and recent tokens", *extra])
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-3000:]
    assert "GENERATION:" in r.stdout and "KV CACHE STATISTICS:" in r.stdout
    import re
    m = re.search(r"Decode Tokens: (\d+)\.", r.stdout)   # 8 unless the random model emits the tokenizer's eos id early
    assert m and 1 <= int(m.group(1)) <= 8, r.stdout[-1500:]
