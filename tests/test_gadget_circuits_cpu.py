"""CPU half of tests/test_gpu_gadget_circuits.py: the same table of gadget circuits (modelled on the
reference's tests/*.rs) through the two composers and the C++ restatement of the prover - layouts
independent of the witness, native export == oracle export, satisfied assignments prove,
unsatisfied ones are refused with CircuitUnsatisfied exactly where the row checker says so."""
import pytest

from oracle import cref
from oracle import gadgets as G
from oracle import pyref as R
from plonk_b200 import gadgets as N
from tests.test_gpu_gadget_circuits import CASES


@pytest.mark.parametrize("name,build,default,satisfied,unsatisfied", CASES, ids=[c[0] for c in CASES])
def test_gadget_circuit_on_cpu(name, build, default, satisfied, unsatisfied):
    def both(vals):
        o, n = G.GadgetComposer.initialized(), N.Composer.initialized()
        build(o, *vals)
        build(n, *vals)
        return o, cref.CircuitArrays(o), n.arrays()

    o, oarr, narr = both(default)
    assert G.unsatisfied_rows(o) == []
    n = 1 << (oarr.constraints + 6 - 1).bit_length()
    cpu = cref.CrefProver(name.encode(), oarr, cref.srs_from_secret(n + 7, 0x5EED + len(name), 0xACE))
    for k, vals in enumerate([default] + satisfied):
        o, oa, na = both(vals)
        assert (na.selectors, na.wires) == (oarr.selectors, oarr.wires), "gate layout must not depend on the witness"
        assert (na.witnesses, na.pi_idx, na.pi_vals) == (oa.witnesses, oa.pi_idx, oa.pi_vals)
        assert G.unsatisfied_rows(o) == [], vals
        assert len(cpu.prove(cref.draw_blinders(R.StdRng.seed_from_u64(100 + k)), oa)) == 1008
    for k, vals in enumerate(unsatisfied):
        o, oa, na = both(vals)
        assert (na.selectors, na.wires) == (oarr.selectors, oarr.wires)
        assert G.unsatisfied_rows(o) != [], vals
        with pytest.raises(ValueError, match="-5"):
            cpu.prove(cref.draw_blinders(R.StdRng.seed_from_u64(200 + k)), oa)
