"""Pins the Python oracle against the reference's own golden vector and Appendix-D checkpoints.

Reference: src/compiler/prover.rs:1132-1162 (deterministic_v3_proof_matches_base_digest)."""
import hashlib

from oracle import pyref as R


def test_stdrng_seed_and_keystream():
    rng = R.StdRng.seed_from_u64(0x9235E700)
    assert [f"{w:08x}" for w in rng.key] == "ecb7c603 5396b474 a97f1681 82ece6ea 777c3f9c df88e9c4 c2652cb1 b7cd4dbe".split()
    assert rng.fill_bytes(16).hex() == "73ad609d1924c8a39ae5886412ec43ec"


def test_merlin_test_vector():
    t = R.Transcript(b"test protocol")
    t.append_message(b"some label", b"some data")
    assert t.challenge_bytes(b"challenge", 32).hex() == "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"


def test_keccak_matches_hashlib_sha3():
    # sha3-256 of the empty string through our permutation
    st = bytearray(200)
    st[0] ^= 0x06
    st[135] ^= 0x80
    R.keccak_f1600(st)
    assert bytes(st[:32]) == hashlib.sha3_256(b"").digest()


def test_constants():
    assert R.fr_to_mont_bytes(R.R_MOD - 1) == b"".join(
        x.to_bytes(8, "little") for x in (0xFFFFFFFD00000003, 0xFB38EC08FFFB13FC, 0x99AD88181CE5880F, 0x5BC8F5F97CD877D8)
    )  # src/composer.rs:334-339
    assert R.ROOT_OF_UNITY == 0x16A2A19EDFE81F20D09B681922C813B4B63683508C2280B93829971F439F0D2B
    assert pow(R.ROOT_OF_UNITY, 1 << 32, R.R_MOD) == 1 and pow(R.ROOT_OF_UNITY, 1 << 31, R.R_MOD) != 1
    assert R.g1_is_on_curve(R.G1_GEN) and R.g1_mul(R.G1_GEN, R.R_MOD - 1) == R.g1_neg(R.G1_GEN)
    assert R.EDWARDS_D == 0x2A9318E74BFA2B48F5FD9207E6BD7FD4292D7F6D37579D2601065FD6D6343EB1


def test_reference_golden_digest_and_checkpoints():
    tr = R.ProofTrace()
    proof = R.kat_proof(tr)
    assert len(proof) == 1008
    assert hashlib.blake2b(proof).digest() == R.KAT_DIGEST
    v = tr.values
    # SURVEY.md Appendix D checkpoints
    assert v["beta"] == 0x536B831CCD8CD794538DEDF8ADDC79C00AEF1D9DAB10A2758EC7F8EC1489CED5
    assert v["gamma"] == 0x0FC9720C42A32D0D9E39C500803B8E89A1B4ED747CD743440635EE5D5AE7BCFC
    assert v["alpha"] == 0x31934A31F92886D23E2715582E68BF4386A930900C9AB96A6A1237B07FC75505
    assert v["z_challenge"] == 0x18D6270C19782C57D7E895EC88E3F87805062C64D84A7696C8C77D84325F3683
    assert R.g1_compress(v["wire_comms"][0]).hex().startswith("88a2541ef26e50d0")
    assert R.g1_compress(v["w_zw_comm"]).hex().startswith("afefc3308740f9e2")
    assert len(v["t_poly"]) == 39 and len(v["r_poly"]) == 15 and len(v["w_z"]) == 14 and len(v["w_zw"]) == 10


def test_commitment_equals_poly_at_secret():
    # independent of pairings: commit(p) == [p(x)] g with the toxic waste known (SURVEY Appendix D note)
    rng = R.StdRng.seed_from_u64(0x9235E700)
    x = R.random_nonzero_bls_scalar(rng)
    gs = R.random_nonzero_bls_scalar(rng)
    pp = R.srs_from_secret(23, x, gs)
    tr = R.ProofTrace()
    R.kat_proof(tr)
    a_poly = tr.values["wire_polys"][0]
    assert R.commit(pp, a_poly) == R.g1_mul(R.g1_mul(R.G1_GEN, gs), R.poly_eval(a_poly, x))
    assert tr.values["wire_comms"][0] == R.commit(pp, a_poly)
