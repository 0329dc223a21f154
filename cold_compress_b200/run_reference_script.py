"""Run one of the reference's own entry scripts (e.g. its unmodified generate.py) against this package:

    python -m cold_compress_b200.run_reference_script /path/to/cold-compress/generate.py --cache_config heavy_hitter ...

The drop-in aliases in cold_compress_b200/dropin/ are put first on sys.path, so the script's
`from cache import ...`, `from generation_utils import ...`, `from tp import maybe_init_dist` bind to
the B200-native implementations; the script's own directory comes next, so everything that is not part of the hot path
(its tokenizer.py, prompts/) stays the reference's.  The script file itself is executed as is.
"""
import runpy
import sys
from pathlib import Path


def main():
    if len(sys.argv) < 2:
        raise SystemExit(__doc__)
    script = str(Path(sys.argv[1]).resolve())
    here = Path(__file__).resolve().parent
    sys.path.insert(0, str(Path(script).parent))
    sys.path.insert(0, str(here.parent))
    sys.path.insert(0, str(here / "dropin"))
    sys.argv = [script] + sys.argv[2:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
