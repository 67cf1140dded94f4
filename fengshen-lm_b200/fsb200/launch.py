"""`python -m fsb200.launch <reference example script> [its arguments ...]` — run an UNMODIFIED Fengshen example on fsb200.

What it does before handing over to the script (runpy, `__main__`):
  * puts fengshen-lm_b200/compat (the `fengshen`, `pytorch_lightning`, `deepspeed` import surfaces of the hot path) and the
    script's own directory in front of sys.path — the reference scripts import siblings such as `llama_generate`;
  * `fsb200.hf.install()`: the HF class names the script imports from `transformers` resolve to fsb200-backed classes.
Under torchrun every rank runs this same line (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* are read by the compat Trainer)."""
import os
import runpy
import sys


def prepare(script):
    here = os.path.dirname(os.path.abspath(__file__))
    compat = os.path.normpath(os.path.join(here, "..", "compat"))
    for p in (os.path.dirname(os.path.abspath(script)), compat, os.path.normpath(os.path.join(here, ".."))):
        if p not in sys.path:
            sys.path.insert(0, p)
    from . import hf
    hf.install()


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] in ("-h", "--help"):
        print(__doc__)
        return 2
    script = argv[0]
    if not os.path.isfile(script):
        raise SystemExit(f"fsb200.launch: no such script: {script}")
    prepare(script)
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")
    return 0


if __name__ == "__main__":
    sys.exit(main())
