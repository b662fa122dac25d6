"""CPU: the reference's shipped example decks through `lmp_b200 -sf b200` with the oracle standing in for the CUDA
library behind the C-ABI (tests/shipped.py), against the unmodified `lmp_serial` on the same text."""
import os

import pytest

import shipped
from shipped import Shipped

pytestmark = pytest.mark.skipif(not shipped.available(), reason="needs /root/reference, oracle/_ref/lmp_serial and lmp_b200 (built by __graft_entry__.build() where the reference exists)")

CASES = shipped.CASES


@pytest.mark.parametrize("case", CASES, ids=[c.name for c in CASES])
def test_shipped_deck_through_the_shells(case, tmp_path):
    shim = shipped.build_shim()
    out = {}
    for who, exe, pre in (("ref", shipped.REF, None), ("b200", shipped.B200, shim)):
        wd = str(tmp_path / who)
        p = shipped.run_one(case, exe, wd, pre)
        assert p.returncode == 0 and "ERROR" not in p.stdout, who + ":\n" + p.stdout[-3000:] + p.stderr[-2000:]
        out[who] = (wd, p.stdout)
    assert "B200 engine: sph_oracle" in out["b200"][1]          # the shells ran, behind them the stand-in
    assert "B200 engine" not in out["ref"][1]
    shipped.compare_rows(shipped.thermo_block(out["ref"][1]), shipped.thermo_block(out["b200"][1]), case.tol, case.name + " thermo")
    files = list(case.files) + (["zz.dump"] if case.dump else [])
    for f in files:
        a = shipped.numeric_rows(os.path.join(out["ref"][0], f))
        b = shipped.numeric_rows(os.path.join(out["b200"][0], f))
        assert len(a) > 0, f
        shipped.compare_rows(a, b, case.tol, case.name + " " + f, shipped.DUMP_VECTORS if f == "zz.dump" else ())
