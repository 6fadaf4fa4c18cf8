"""tests/models/quotient_4n_model.py: the quotient polynomial recovered from a 4n coset plus eight extra
points equals the oracle's (8n coset, src/proof_system/quotient_poly.rs:20-137), and a corrupted
witness is rejected by both.  Design validation for the next round-3 kernel schedule."""
import importlib.util
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def _model():
    spec = importlib.util.spec_from_file_location("quotient_4n_model", os.path.join(HERE, "models", "quotient_4n_model.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_quotient_from_4n_coset_matches_reference_and_rejects_bad_witness():
    m = _model()
    m.check(20, 1, 0)
    m.check(60, 2, 7)
    m.check(60, 4, 7, corrupt=True)


def test_lagrange_basis_wire_commitments_equal_commit_of_blinded_polynomials():
    """tests/models/lagrange_commit_model.py: design validation for committing to the wire polynomials
    through their values (short scalars) instead of their coefficients."""
    spec = importlib.util.spec_from_file_location("lagrange_commit_model", os.path.join(HERE, "models", "lagrange_commit_model.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    m.check(16, 3)
    m.check(8, 5)


def test_batched_affine_tree_layouts_and_work_split():
    """tests/models/affine_tree_model.py: the position-based work split and the halving layouts of the opt-in
    batched-affine bucket accumulation stay inside their buffers and add every bucket exactly once."""
    spec = importlib.util.spec_from_file_location("affine_tree_model", os.path.join(HERE, "models", "affine_tree_model.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    m.check()
