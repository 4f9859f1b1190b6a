"""The import paths the reference scripts use (inference.py:8-12, ex_audioset.py:16-22) resolve to this
repository's modules when `dropin/` is on sys.path."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_import_paths_resolve_to_this_package():
    code = (
        "from models.mn.model import get_model as get_mobilenet\n"
        "from models.dymn.model import get_model as get_dymn\n"
        "from models.preprocess import AugmentMelSTFT\n"
        "from helpers.utils import NAME_TO_WIDTH, exp_warmup_linear_down, mixup\n"
        "import models.mn.model as m\n"
        "assert 'efficientat_b200' in get_mobilenet.__module__ and 'efficientat_b200' in AugmentMelSTFT.__module__\n"
        "assert NAME_TO_WIDTH('mn10_as') == 1.0 and NAME_TO_WIDTH('dymn20_as(2)') == 2.0 and NAME_TO_WIDTH('x') == 1.0\n"
        "net = get_mobilenet(width_mult=NAME_TO_WIDTH('mn04_as'), verbose=False)\n"
        "assert sum(p.numel() for p in net.parameters()) == 983599, sum(p.numel() for p in net.parameters())\n"
        "print('ok')\n")
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "dropin") + os.pathsep + ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=ROOT)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]


def test_param_counts_match_reference_table():
    """README.md:94-113 / SURVEY.md section 4: 4.88 M (mn10), 10.55 M (dymn10) parameters."""
    from efficientat_b200.models.dymn.model import get_model as dymn
    from efficientat_b200.models.mn.model import get_model as mn
    assert sum(p.numel() for p in mn(verbose=False).parameters()) == 4876831
    assert sum(p.numel() for p in dymn(verbose=False).parameters()) == 10548479
