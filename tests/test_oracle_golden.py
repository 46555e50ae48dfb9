"""CPU: the C oracle (oracle/tc_oracle.c) against the golden vectors produced by the independent
big-int model (oracle/pyref.py, tests/golden/gen_golden.py) and against the two public anchors."""
import numpy as np

from conftest import fr_bytes, hx, hxs, R


def test_public_anchors(O, golden):
    a = golden["anchors"]
    assert bytes(O.g1_compress(O.g1_generator())[0]).hex() == a["g1_gen_compressed"]
    two = O.g1_mul_gen_batch(fr_bytes([2]))
    assert bytes(two[0]).hex() == a["g1_times_2_uncompressed"]


def test_sha3_and_chacha(O, golden):
    import ctypes as C
    for m, d in golden["sha3_256"]:
        out = np.zeros(32, np.uint8)
        mb = np.frombuffer(bytes.fromhex(m) or b"\0", np.uint8).copy()
        O.lib().orc_sha3_256(mb.ctypes.data_as(C.c_void_p), C.c_size_t(len(m) // 2), out.ctypes.data_as(C.c_void_p))
        assert bytes(out).hex() == d
    # ChaCha20 zero-key block 0 (RFC 7539 test vector first words 0xade0b876, 0x903df1a0, ...)
    assert golden["chacha_zero_key_words"][:2] == [0xade0b876, 0x903df1a0]
    fr = golden["fr_random_stream"]
    got = O.fr_random_stream(bytes.fromhex(fr["seed"]), len(fr["out"]))
    assert [bytes(r).hex() for r in got] == fr["out"]


def test_keys_hashes_signatures(O, golden):
    sk = np.concatenate([hx(s) for s in golden["sk"]])
    pk = O.g1_mul_gen_batch(sk)
    assert [bytes(p).hex() for p in pk] == golden["pk"]
    assert [bytes(p).hex() for p in O.g1_compress(pk)] == golden["pk_compressed"]
    msgs = [bytes.fromhex(m) for m in golden["msgs"]]
    h = O.hash_g2_batch(msgs)
    assert [bytes(p).hex() for p in h] == golden["hash_g2"]
    assert [bytes(p).hex() for p in O.g2_compress(h)] == golden["hash_g2_compressed"]
    nsk = len(golden["sk"])
    sks = np.concatenate([hx(golden["sk"][i % nsk]) for i in range(len(msgs))])
    assert [bytes(s).hex() for s in O.sign_batch(sks, msgs)] == golden["sig"]
    assert [bytes(s).hex() for s in O.sign_g2_batch(sks, h)] == golden["sig"]


def test_verify_cases(O, golden):
    msgs = [bytes.fromhex(m) for m in golden["msgs"]]
    for pi, si, mi, exp in golden["verify_cases"]:
        ok = O.verify_batch(hx(golden["pk"][pi]), hx(golden["sig"][si]), [msgs[mi]])
        assert bool(ok[0]) == exp
        ok = O.verify_g2_batch(hx(golden["pk"][pi]), hx(golden["hash_g2"][mi]), None, hx(golden["sig"][si]))
        assert bool(ok[0]) == exp


def test_gt_is_cube_of_textbook_pairing(O):
    """The x-chain hard part (as recalled from pairing 0.16) yields e(P,Q)^3; GT is unobservable
    through the API, the relation is recorded here so the oracle's GT convention is pinned."""
    import pyref as Y
    gt = bytes(O.pairing_gt(O.g1_generator(), O.g2_generator()))
    vals = [int.from_bytes(gt[48 * i:48 * i + 48], "big") for i in range(12)]
    g = [None] * 6
    for idx, m in enumerate([0, 2, 4, 1, 3, 5]):
        g[m] = (vals[2 * idx], vals[2 * idx + 1])
    e = Y.pairing(Y.G1_GEN, Y.G2_GEN)
    assert g == Y.f12_mul(Y.f12_mul(e, e), e)


def test_threshold_sig(O, golden):
    ts = golden["threshold_sig"]
    t = ts["t"]
    for s in ts["sets"]:
        x = fr_bytes([i + 1 for i in s["idx"]])
        out, st = O.combine_g2_batch(1, t, x, hxs(s["shares"]))
        assert st[0] == 0 and bytes(out[0]).hex() == s["combined"]
    comm = hxs(golden["commitment"])
    shares = O.commitment_eval_batch(comm, fr_bytes([i + 1 for i in range(5)]))
    assert [bytes(p).hex() for p in shares] == golden["pk_shares"]
    # master public key verifies the combined signature
    ok = O.verify_batch(comm[0], hx(ts["sets"][0]["combined"]), [bytes.fromhex(ts["msg"])])
    assert ok[0] == 1


def test_threshold_enc(O, golden):
    e = golden["enc"]
    pk = hx(golden["commitment"][0])
    u, v, w = O.encrypt(pk, hx(e["r"]), bytes.fromhex(e["plain"]))
    assert bytes(u).hex() == e["u"] and v.hex() == e["v"] and bytes(w).hex() == e["w"]
    assert bytes(O.hash_g1_g2(u, v)).hex() == e["hash_g1_g2"]
    # Ciphertext::verify: e(g1, W) == e(U, H(U,V))
    assert O.verify_g2_batch(O.g1_generator(), w, u, O.hash_g1_g2(u, v))[0] == 1
    poly = np.concatenate([hx(c) for c in golden["poly"]])
    ski = O.poly_eval(poly, fr_bytes([i + 1 for i in e["idx"]]))
    ds = O.decrypt_share_batch(ski, np.tile(u, (len(e["idx"]), 1)))
    assert [bytes(d).hex() for d in ds] == e["dshares"]
    out, st = O.decrypt_batch(1, 2, fr_bytes([i + 1 for i in e["idx"]]), ds, [v])
    assert st[0] == 0 and out[0].hex() == e["decrypted"] == e["plain"]
    hl = golden["hash_g1_g2_long"]
    assert bytes(O.hash_g1_g2(hx(hl["g1"]), bytes.fromhex(hl["msg"]))).hex() == hl["out"]
    xw = golden["xor_with_hash"]
    assert O.xor_with_hash(hx(xw["g1"]), bytes.fromhex(xw["in"])).hex() == xw["out"]


def test_commit_eval_and_poly_kat(O, golden):
    ce = golden["commit_eval"]
    out = O.commitment_eval_batch(hxs(ce["coeff"]), np.concatenate([hx(x) for x in ce["x"]]))
    assert [bytes(p).hex() for p in out] == ce["out"]
    pk = golden["poly_kat"]   # reference's own KAT, src/poly.rs:783-797
    coeff = np.concatenate([hx(c) for c in pk["coeff"]])
    for x, y in pk["samples"]:
        assert bytes(O.poly_eval(coeff, hx(x))[0]).hex() == y


def test_decoders_reject_bad_encodings(O, golden):
    bad = hxs(golden["g1_bad_compressed"])
    _, st = O.g1_decompress(bad)
    assert list(st) == [3] * len(bad)
    good = hxs(golden["pk_compressed"])
    unc, st = O.g1_decompress(good)
    assert list(st) == [0] * len(good) and [bytes(p).hex() for p in unc] == golden["pk"]
    unc, st = O.g2_decompress(hxs(golden["hash_g2_compressed"]))
    assert list(st) == [0] * len(unc) and [bytes(p).hex() for p in unc] == golden["hash_g2"]
    unc, st = O.g1_decompress(hx(golden["g1_inf_compressed"]))
    assert st[0] == 0 and unc[0][0] == 0x40
    unc, st = O.g2_decompress(hx(golden["g2_inf_compressed"]))
    assert st[0] == 0 and unc[0][0] == 0x40


def test_infinity_operands(O):
    """A pairing with an infinity operand is 1."""
    inf1 = np.zeros(96, np.uint8); inf1[0] = 0x40
    inf2 = np.zeros(192, np.uint8); inf2[0] = 0x40
    g1, g2 = O.g1_generator(), O.g2_generator()
    assert O.verify_g2_batch(inf1, g2, g1, inf2)[0] == 1
    assert O.verify_g2_batch(g1, g2, g1, inf2)[0] == 0
