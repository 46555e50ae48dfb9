// The reference's scheme tests (src/lib.rs:811-939) written against the C++ mirror include/tcb200.hpp.
// Linked against libtcb200.so on the GPU box and against the host-emulation build (test infrastructure)
// on the CPU box — same symbols, same source.
#include <cstdio>
#include <cstdlib>
#include "../../include/tcb200.hpp"
using namespace tcb200;

static uint64_t rs = 0x9e3779b97f4a7c15ULL;
static uint64_t rnd() { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return rs; }
static Fr rand_fr() { Fr f; for (int i = 0; i < 4; i++) f.l[i] = rnd(); f.l[3] >>= 2; return f; }   // < 2^254 < r
static Bytes B(const char *s) { return Bytes(s, s + strlen(s)); }
#define CHECK(c) do { if (!(c)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

static int test_simple_sig(Engine &e) {            // src/lib.rs:811-820
    SecretKey sk0{rand_fr().bytes()}, sk1{rand_fr().bytes()};
    PublicKey pk0 = sk0.public_key(e);
    Bytes msg0 = B("Real news"), msg1 = B("Fake news");
    CHECK(pk0.verify(e, sk0.sign(e, msg0), msg0));
    CHECK(!pk0.verify(e, sk1.sign(e, msg0), msg0));
    CHECK(!pk0.verify(e, sk0.sign(e, msg1), msg0));
    return 0;
}
static int test_threshold_sig(Engine &e) {         // src/lib.rs:823-873
    SecretKeySet sk_set;
    for (int i = 0; i < 4; i++) sk_set.poly.coeff.push_back(rand_fr());     // threshold 3
    PublicKeySet pk_set = sk_set.public_keys(e);
    Bytes msg = B("Totally real news");
    std::map<uint64_t, SignatureShare> sigs;
    for (uint64_t i : {5, 8, 7, 10}) sigs[i] = sk_set.secret_key_share(i).sign_share(e, msg);
    for (auto &kv : sigs) CHECK(pk_set.public_key_share(e, kv.first).verify(e, kv.second, msg));
    Signature sig = pk_set.combine_signatures(e, sigs);
    CHECK(pk_set.public_key().verify(e, sig, msg));
    std::map<uint64_t, SignatureShare> sigs2;
    for (uint64_t i : {42, 43, 44, 45}) sigs2[i] = sk_set.secret_key_share(i).sign_share(e, msg);
    CHECK(pk_set.combine_signatures(e, sigs2) == sig);                       // same signature from disjoint shares
    sigs.erase(10);
    bool threw = false;
    try { pk_set.combine_signatures(e, sigs); } catch (const NotEnoughShares &) { threw = true; }
    CHECK(threw);
    return 0;
}
static int test_threshold_enc(Engine &e) {         // src/lib.rs:876-939
    SecretKeySet sk_set;
    for (int i = 0; i < 3; i++) sk_set.poly.coeff.push_back(rand_fr());     // threshold 2
    PublicKeySet pk_set = sk_set.public_keys(e);
    Bytes msg = B("Totally real news");
    Ciphertext ct = pk_set.public_key().encrypt_with_r(e, rand_fr().bytes(), msg);
    CHECK(ciphertext_verify(e, ct));
    std::map<uint64_t, DecryptionShare> shares;
    for (uint64_t i : {8, 5, 9}) {
        SecretKeyShare s = sk_set.secret_key_share(i);
        shares[i] = s.decrypt_share_no_verify(e, ct);
        CHECK(pk_set.public_key_share(e, i).verify_decryption_share(e, shares[i], ct));
    }
    CHECK(pk_set.decrypt(e, shares, ct) == msg);
    SecretKey master{sk_set.poly.evaluate(Fr::from_u64(0)).bytes()};
    auto dec = master.decrypt(e, ct);
    CHECK(dec.first && dec.second == msg);
    Ciphertext fake = ct; fake.v[0] ^= 1;
    CHECK(!ciphertext_verify(e, fake) && !master.decrypt(e, fake).first);
    std::vector<PublicKey> pks{pk_set.public_key(), pk_set.public_key()};
    std::vector<Signature> sg{master.sign(e, msg), master.sign(e, B("x"))};
    std::vector<bool> ok = verify_batch(e, pks, sg, {msg, msg});
    CHECK(ok[0] && !ok[1]);
    return 0;
}
int main() {
    Engine e;
    int rc = test_simple_sig(e) | test_threshold_sig(e) | test_threshold_enc(e);
    std::printf(rc ? "cpp mirror tests FAILED\n" : "cpp mirror tests ok\n");
    return rc;
}
