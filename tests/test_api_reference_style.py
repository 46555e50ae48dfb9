"""The reference's own scheme tests (src/lib.rs:790-940, src/poly.rs:783-797), re-expressed on the
Python mirror of its API (threshold_crypto_b200/api.py).  On the CPU box they run on the
host-emulation build of the device code (logic check); on the B200 they run on the CUDA library."""
import numpy as np
import pytest

from threshold_crypto_b200 import api


@pytest.fixture(params=["emu", pytest.param("gpu", marks=pytest.mark.gpu)])
def tc(request):
    if request.param == "emu":
        api.set_engine(request.getfixturevalue("emu"))
    else:
        api.set_engine(request.getfixturevalue("gpu_engine"))
    yield api
    api.set_engine(None)


def rng():
    return np.random.default_rng(20260925)


def test_simple_sig(tc):                      # src/lib.rs:811-820
    r = rng()
    sk0, sk1 = tc.SecretKey(int.from_bytes(r.bytes(40), "little")), tc.SecretKey(int.from_bytes(r.bytes(40), "little"))
    pk0 = sk0.public_key()
    msg0, msg1 = b"Real news", b"Fake news"
    assert pk0.verify(sk0.sign(msg0), msg0)
    assert not pk0.verify(sk1.sign(msg0), msg0)     # wrong key
    assert not pk0.verify(sk0.sign(msg1), msg0)     # wrong message


def test_threshold_sig(tc):                   # src/lib.rs:823-873
    sk_set = tc.SecretKeySet.random(3, rng())
    pk_set = sk_set.public_keys()
    pk_master = pk_set.public_key()
    assert pk_master != pk_set.public_key_share(0)
    assert pk_master == sk_set.secret_key().public_key()
    msg = b"Totally real news"
    sigs = {i: sk_set.secret_key_share(i).sign(msg) for i in (5, 8, 7, 10)}
    for i, s in sigs.items():
        assert pk_set.public_key_share(i).verify(s, msg)
    sig = pk_set.combine_signatures(sigs)
    assert pk_set.public_key().verify(sig, msg)
    sigs2 = {i: sk_set.secret_key_share(i).sign(msg) for i in (42, 43, 44, 45)}
    sig2 = pk_set.combine_signatures(sigs2)
    assert sig == sig2                               # two disjoint share sets give the SAME signature
    with pytest.raises(tc.NotEnoughShares):
        pk_set.combine_signatures({5: sigs[5], 8: sigs[8], 7: sigs[7]})


def test_interpolate_equals_evaluate_at_zero(tc):   # src/lib.rs:794-808
    r = rng()
    for deg in range(0, 4):
        poly = tc.Poly.random(deg, r)
        comm = poly.commitment()
        x, vals = 0, []
        for _ in range(deg + 1):
            x += int(r.integers(1, 5))
            vals.append((x - 1, tc.DecryptionShare(comm.evaluate(x))))
        # interpolate over G1 == comm.evaluate(0); go through PublicKeySet.decrypt's machinery with an empty body
        pts = np.stack([v.raw for _, v in vals])
        xs = tc._frs([tc.into_fr_plus_1(i) for i, _ in vals])
        out, st = tc.engine().combine_g1_batch(1, deg, xs, pts)
        assert not st.any() and np.array_equal(out[0], comm.evaluate(0))


def test_threshold_enc_decrypt_roundtrip(tc, O):     # src/lib.rs:908-939 (encrypt done by the oracle: next-row item)
    sk_set = tc.SecretKeySet.random(2, rng())
    pk_set = sk_set.public_keys()
    msg = b"Totally real news"
    r32 = tc._fr(0x1234567890abcdef1234567890abcdef)
    u, v, w = O.encrypt(pk_set.public_key().raw, r32, msg)
    ct = tc.Ciphertext(u, v, w)
    assert ct.verify()                                            # src/lib.rs:508-512
    assert not tc.Ciphertext(u, v + b"x", w).verify()             # tampered body (src/lib.rs:893-896)
    shares = {i: sk_set.secret_key_share(i).decrypt_share_no_verify(ct) for i in (8, 5, 9)}
    for i, s in shares.items():
        assert pk_set.public_key_share(i).verify_decryption_share(s, ct)
    assert not pk_set.public_key_share(5).verify_decryption_share(shares[8], ct)
    assert pk_set.decrypt(shares, ct) == msg
    with pytest.raises(tc.NotEnoughShares):
        pk_set.decrypt({8: shares[8], 5: shares[5]}, ct)


def test_hash_g2(tc, O):                      # src/lib.rs:943-953, plus the bytes of the oracle (1000-byte messages: several SHA3 blocks)
    r = rng()
    msg = r.bytes(1000)
    msgs = [msg, msg, msg + b"end0", msg + b"end1"]
    h = tc.engine().hash_g2_batch(msgs)
    assert np.array_equal(h[0], h[1])
    assert not np.array_equal(h[0], h[2]) and not np.array_equal(h[2], h[3])
    assert np.array_equal(h, O.hash_g2_batch(msgs))


def test_hash_g1_g2(tc, O):                   # src/lib.rs:956-969 (messages longer than 64 bytes are hashed first)
    r = rng()
    msg = r.bytes(1000)
    g0, g1 = tc.SecretKey(int.from_bytes(r.bytes(40), "little")).public_key().raw, tc.SecretKey(int.from_bytes(r.bytes(40), "little")).public_key().raw
    pts = np.stack([g0, g0, g0, g0, g1])
    msgs = [msg, msg, msg + b"end0", msg + b"end1", msg]
    h = tc.engine().hash_g1_g2_batch(pts, msgs)
    assert np.array_equal(h[0], h[1])
    assert not np.array_equal(h[0], h[2]) and not np.array_equal(h[2], h[3]) and not np.array_equal(h[0], h[4])
    for k in range(5):
        assert np.array_equal(h[k], O.hash_g1_g2(pts[k], msgs[k]))


def test_xor_with_hash(tc, O):                # src/lib.rs:972-982, through decrypt with t = 0 (the body is xored with the hash of the share)
    r = rng()
    g0, g1 = tc.SecretKey(int.from_bytes(r.bytes(40), "little")).public_key().raw, tc.SecretKey(int.from_bytes(r.bytes(40), "little")).public_key().raw
    E = tc.engine()
    x1 = tc._frs([1])

    def xwh(g, body):
        out, st = E.decrypt_batch(1, 0, x1, g, [body])
        assert not st.any()
        return out[0]
    assert xwh(g0, bytes(5)) == xwh(g0, bytes(5)) and xwh(g0, bytes(5)) != xwh(g1, bytes(5))
    for n in (5, 6, 20):
        assert len(xwh(g0, bytes(n))) == n
    assert xwh(g0, bytes(20)) == O.decrypt_batch(1, 0, x1, g0, [bytes(20)])[0][0]


def test_random_extreme_thresholds(tc):       # src/lib.rs:900-905: threshold 0 — every share is the master key
    sks = tc.SecretKeySet.random(0, rng())
    assert sks.threshold() == 0
    pks = sks.public_keys()
    msg = b"one share suffices"
    share = sks.secret_key_share(7).sign(msg)
    assert pks.public_key_share(7).verify(share, msg)
    assert pks.combine_signatures({7: share}) == sks.secret_key().sign(msg)
    assert pks.public_key().verify(pks.combine_signatures({7: share}), msg)


def test_poly_kat(tc):                        # src/poly.rs:783-797: 5 X^3 + X - 2
    poly = tc.Poly([-2, 1, 0, 5])
    for x, y in ((-1, -8), (2, 40), (3, 136), (5, 628)):
        assert poly.evaluate(x) == y % tc.R


def test_simple_enc(tc):                       # src/lib.rs:876-897
    r = rng()
    sk_bob, sk_eve = tc.SecretKey(int.from_bytes(r.bytes(40), "little")), tc.SecretKey(int.from_bytes(r.bytes(40), "little"))
    pk_bob = sk_bob.public_key()
    msg = b"Muffins in the canteen today! Don't tell Eve!"
    ct = pk_bob.encrypt_with_rng(r, msg)
    assert ct.verify()
    assert sk_bob.decrypt(ct) == msg
    assert sk_eve.decrypt(ct) != msg                      # Eve gets garbage
    fake = tc.Ciphertext(ct.u, b"fake news" + ct.v[9:], ct.w)
    assert not fake.verify() and sk_bob.decrypt(fake) is None


def test_distributed_key_generation(tc):       # src/poly.rs:819-900 (dealers 3, nodes 5, faulty 2)
    r = rng()
    dealer_num, node_num, faulty_num = 3, 5, 2
    sec_keys = [0] * node_num
    pub_bivar_polys = []
    for _ in range(dealer_num):
        bi_poly = tc.BivarPoly.random(faulty_num, r)
        bi_commit = bi_poly.commitment()
        pub_bivar_polys.append(bi_commit)
        for m in range(1, node_num + 1):
            row_poly = bi_poly.row(m)
            row_commit = bi_commit.row(m)
            assert row_poly.commitment() == row_commit                      # the row matches the public commitment
            for s in range(1, node_num + 1):
                val = row_poly.evaluate(s)
                assert np.array_equal(bi_commit.evaluate(m, s), tc.engine().g1_mul_gen_batch(tc._fr(val))[0])
                assert bi_poly.evaluate(m, s) == val
            # f(m, 0) from faulty_num + 1 column values by interpolation (src/poly.rs:868-877)
            received = [(s, bi_poly.evaluate(m, s)) for s in range(1, faulty_num + 2)]
            sec_keys[m - 1] = (sec_keys[m - 1] + tc.Poly.interpolate(received).evaluate(0)) % tc.R
    # the summed commitment's row(0) is the master commitment; its evaluations are the key shares' public keys
    sum_commit = pub_bivar_polys[0].row(0)
    for bc in pub_bivar_polys[1:]:
        sum_commit = sum_commit + bc.row(0)
    for m in range(1, node_num + 1):
        assert np.array_equal(sum_commit.evaluate(m), tc.engine().g1_mul_gen_batch(tc._fr(sec_keys[m - 1]))[0])


def test_poly_algebra(tc):                    # src/poly.rs:783-797 incl. interpolation
    x3, x1 = tc.Poly.monomial(3), tc.Poly.monomial(1)
    poly = x3 * 5 + x1 - 2
    assert poly == tc.Poly([-2, 1, 0, 5])
    samples = [(-1, -8), (2, 40), (3, 136), (5, 628)]
    for x, y in samples:
        assert poly.evaluate(x) == y % tc.R
    assert tc.Poly.interpolate(samples) == poly


def test_checked_from_bytes_and_size_errors(tc):
    """PublicKey / Signature can only come from the checked decoder (src/lib.rs:140-146, 246-252): round trip through the
    48 / 96-byte wire format, FromBytesError for a tampered encoding and for wrong lengths; wrong buffer sizes raise before the
    C ABI is entered."""
    A = tc
    r = rng()
    sk = A.SecretKey(int.from_bytes(r.bytes(40), "little"))
    pk, sig = sk.public_key(), sk.sign(b"wire format")
    assert len(pk.to_bytes()) == A.PK_SIZE and len(sig.to_bytes()) == A.SIG_SIZE
    assert A.PublicKey.from_bytes(pk.to_bytes()) == pk and A.Signature.from_bytes(sig.to_bytes()) == sig
    bad = bytearray(pk.to_bytes()); bad[47] ^= 1
    with pytest.raises(A.FromBytesError):
        A.PublicKey.from_bytes(bytes(bad))
    with pytest.raises(A.FromBytesError):
        A.Signature.from_bytes(b"\x00" * 95)
    with pytest.raises(ValueError):
        A.PublicKey(np.zeros(95, np.uint8))
    with pytest.raises(ValueError):
        A.engine().verify_g2_batch(pk.raw, sig.raw[:100], None, sig.raw)
    with pytest.raises(ValueError):
        A.engine().decrypt_batch(1, 1, np.zeros(64, np.uint8), np.zeros(96, np.uint8), [b"x"])
    assert A.Commitment(np.zeros((0, 96), np.uint8)).evaluate(3)[0] == 0x40          # empty commitment -> G1::zero()
