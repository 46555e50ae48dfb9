// tcb200.hpp — header-only C++ mirror of the reference's scheme API for the hot path, on top of the C ABI
// (tcb200.h).  The reference is a Rust crate and no Rust toolchain exists in the build image, so the host
// side above the ABI is C++ (this file, for compiled callers) and Python (threshold_crypto_b200/api.py, for
// the tests).  Names, argument meaning and error behaviour follow /root/reference/src/lib.rs and
// src/poly.rs (cited per method); all curve arithmetic happens on the GPU through tcb_ctx, only the cheap
// Fr polynomial algebra (Poly::evaluate, index -> i + 1, take(t+1)) stays on the host.
#pragma once
#include <array>
#include <cstdint>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "tcb200.h"

namespace tcb200 {

typedef std::array<uint8_t, 32> FrBytes;    // canonical little-endian Fr (src/serde_impl.rs:109)
typedef std::array<uint8_t, 96> G1Bytes;    // uncompressed affine G1
typedef std::array<uint8_t, 192> G2Bytes;   // uncompressed affine G2

struct Error : std::runtime_error { using std::runtime_error::runtime_error; };          // src/error.rs:7-20
struct NotEnoughShares : Error { NotEnoughShares() : Error("Not enough shares for interpolation") {} };
struct DuplicateEntry : Error { DuplicateEntry() : Error("Samples for interpolation contain duplicate entries") {} };

// ---- minimal Fr arithmetic for the host-side polynomial algebra (255-bit, schoolbook + shift-subtract)
struct Fr {
    uint64_t l[4];
    static const uint64_t *modulus() {
        static const uint64_t r[4] = {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL};
        return r;
    }
    static Fr from_u64(uint64_t v) { return Fr{{v, 0, 0, 0}}; }
    static bool geq(const uint64_t *a, const uint64_t *b) {
        for (int i = 3; i >= 0; i--) { if (a[i] > b[i]) return true; if (a[i] < b[i]) return false; }
        return true;
    }
    Fr add(const Fr &o) const {
        Fr r; unsigned __int128 c = 0;
        for (int i = 0; i < 4; i++) { c += (unsigned __int128)l[i] + o.l[i]; r.l[i] = (uint64_t)c; c >>= 64; }
        if (c || geq(r.l, modulus())) { unsigned __int128 b = 0; for (int i = 0; i < 4; i++) { unsigned __int128 d = (unsigned __int128)r.l[i] - modulus()[i] - (uint64_t)b; r.l[i] = (uint64_t)d; b = (d >> 64) & 1; } }
        return r;
    }
    Fr mul(const Fr &o) const {   // double-and-add on the bits of o (host side, tiny degree)
        Fr acc{{0, 0, 0, 0}};
        for (int i = 255; i >= 0; i--) {
            acc = acc.add(acc);
            if ((o.l[i / 64] >> (i % 64)) & 1) acc = acc.add(*this);
        }
        return acc;
    }
    FrBytes bytes() const { FrBytes b; std::memcpy(b.data(), l, 32); return b; }
    static Fr from_bytes(const FrBytes &b) { Fr r; std::memcpy(r.l, b.data(), 32); return r; }
};
inline FrBytes into_fr_plus_1(uint64_t i) { return Fr::from_u64(i).add(Fr::from_u64(1)).bytes(); }   // src/lib.rs:769-773

// ---- the GPU context (one per calling thread, like every &self method of the reference is re-entrant)
class Engine {
  public:
    explicit Engine(const std::vector<int> &devices = {}) {
        int rc = tcb_init(&ctx_, devices.empty() ? nullptr : devices.data(), (int)devices.size());
        if (rc != 0) throw Error("tcb_init failed: no usable CUDA device (there is no CPU fallback)");
    }
    ~Engine() { tcb_free(ctx_); }
    Engine(const Engine &) = delete;
    Engine &operator=(const Engine &) = delete;
    tcb_ctx *ctx() const { return ctx_; }
    void check(int rc) const { if (rc != 0) throw Error(std::string("tcb call failed: ") + tcb_last_error(ctx_)); }

  private:
    tcb_ctx *ctx_ = nullptr;
};

typedef std::vector<uint8_t> Bytes;
struct Msgs {   // concatenated messages + offsets, the ABI's ragged-batch form
    Bytes buf;
    std::vector<uint64_t> off{0};
    void push(const Bytes &m) { buf.insert(buf.end(), m.begin(), m.end()); off.push_back(buf.size()); }
    const uint8_t *data() const { static const uint8_t z = 0; return buf.empty() ? &z : buf.data(); }
};

struct Signature { G2Bytes raw; bool operator==(const Signature &o) const { return raw == o.raw; } };   // src/lib.rs:202
struct SignatureShare : Signature {};                                                                      // src/lib.rs:266
struct DecryptionShare { G1Bytes raw; };                                                                   // src/lib.rs:517
struct Ciphertext { G1Bytes u; Bytes v; G2Bytes w; };                                                      // src/lib.rs:474-478

struct PublicKey {                                                                                          // src/lib.rs:79
    G1Bytes raw;
    bool operator==(const PublicKey &o) const { return raw == o.raw; }
    bool verify(const Engine &e, const Signature &sig, const Bytes &msg) const {                           // src/lib.rs:115-117
        Msgs m; m.push(msg);
        uint8_t ok = 0;
        e.check(tcb_verify_batch(e.ctx(), 1, raw.data(), sig.raw.data(), m.data(), m.off.data(), &ok));
        return ok != 0;
    }
    Ciphertext encrypt_with_r(const Engine &e, const FrBytes &r, const Bytes &msg) const {                 // src/lib.rs:128-137 (r from the caller's RNG)
        Msgs m; m.push(msg);
        Ciphertext ct; ct.v.resize(msg.size());
        uint8_t dummy = 0;
        e.check(tcb_encrypt_batch(e.ctx(), 1, raw.data(), r.data(), m.data(), m.off.data(), ct.u.data(), msg.empty() ? &dummy : ct.v.data(), ct.w.data()));
        return ct;
    }
};
struct PublicKeyShare : PublicKey {                                                                         // src/lib.rs:160
    bool verify_decryption_share(const Engine &e, const DecryptionShare &share, const Ciphertext &ct) const {   // src/lib.rs:182-186
        Msgs m; m.push(ct.v);
        G2Bytes h;
        e.check(tcb_hash_g1_g2_batch(e.ctx(), 1, ct.u.data(), m.data(), m.off.data(), h.data()));
        uint8_t ok = 0;
        e.check(tcb_verify_g2_batch(e.ctx(), 1, share.raw.data(), h.data(), raw.data(), ct.w.data(), &ok));
        return ok != 0;
    }
};
inline bool ciphertext_verify(const Engine &e, const Ciphertext &ct) {                                     // Ciphertext::verify, src/lib.rs:508-512
    Msgs m; m.push(ct.v);
    G2Bytes h;
    e.check(tcb_hash_g1_g2_batch(e.ctx(), 1, ct.u.data(), m.data(), m.off.data(), h.data()));
    uint8_t ok = 0;
    e.check(tcb_verify_g2_batch(e.ctx(), 1, ct.u.data(), h.data(), nullptr, ct.w.data(), &ok));
    return ok != 0;
}

struct SecretKey {                                                                                          // src/lib.rs:302
    FrBytes fr;
    PublicKey public_key(const Engine &e) const {                                                           // src/lib.rs:367-369
        PublicKey pk;
        e.check(tcb_g1_mul_gen_batch(e.ctx(), 1, fr.data(), pk.raw.data()));
        return pk;
    }
    Signature sign(const Engine &e, const Bytes &msg) const {                                               // src/lib.rs:379-381
        Msgs m; m.push(msg);
        Signature s;
        e.check(tcb_sign_batch(e.ctx(), 1, fr.data(), m.data(), m.off.data(), s.raw.data()));
        return s;
    }
    // None (empty optional-like pair) if the ciphertext is invalid, src/lib.rs:384-391
    std::pair<bool, Bytes> decrypt(const Engine &e, const Ciphertext &ct) const {
        if (!ciphertext_verify(e, ct)) return {false, {}};
        G1Bytes g;
        e.check(tcb_decrypt_share_batch(e.ctx(), 1, fr.data(), ct.u.data(), g.data()));
        Msgs m; m.push(ct.v);
        Bytes out(ct.v.size() ? ct.v.size() : 1);
        uint8_t st = 0;
        FrBytes one = into_fr_plus_1(0);
        e.check(tcb_decrypt_batch(e.ctx(), 1, 0, one.data(), g.data(), m.data(), m.off.data(), out.data(), &st));
        out.resize(ct.v.size());
        return {true, out};
    }
};
struct SecretKeyShare : SecretKey {                                                                         // src/lib.rs:412
    PublicKeyShare public_key_share(const Engine &e) const { PublicKeyShare p; p.raw = public_key(e).raw; return p; }
    SignatureShare sign_share(const Engine &e, const Bytes &msg) const { SignatureShare s; s.raw = sign(e, msg).raw; return s; }
    DecryptionShare decrypt_share_no_verify(const Engine &e, const Ciphertext &ct) const {                  // src/lib.rs:460-462
        DecryptionShare d;
        e.check(tcb_decrypt_share_batch(e.ctx(), 1, fr.data(), ct.u.data(), d.raw.data()));
        return d;
    }
};

struct Poly {                                                                                               // poly::Poly, src/poly.rs:40-44
    std::vector<Fr> coeff;
    size_t degree() const { return coeff.size() - 1; }
    Fr evaluate(const Fr &x) const {                                                                        // src/poly.rs:358-369
        Fr acc{{0, 0, 0, 0}};
        for (size_t k = coeff.size(); k-- > 0;) acc = acc.mul(x).add(coeff[k]);
        return acc;
    }
};
struct Commitment {                                                                                         // poly::Commitment, src/poly.rs:429-433
    std::vector<G1Bytes> coeff;
    size_t degree() const { return coeff.size() - 1; }
    G1Bytes evaluate(const Engine &e, const FrBytes &x) const {                                             // src/poly.rs:497-508
        G1Bytes out;
        e.check(tcb_commitment_eval_batch(e.ctx(), degree(), coeff[0].data(), 1, x.data(), out.data()));
        return out;
    }
};

class PublicKeySet {                                                                                        // src/lib.rs:539-626
  public:
    Commitment commit;
    size_t threshold() const { return commit.degree(); }
    PublicKey public_key() const { PublicKey p; p.raw = commit.coeff[0]; return p; }
    PublicKeyShare public_key_share(const Engine &e, uint64_t i) const {                                    // src/lib.rs:570-573
        PublicKeyShare p; p.raw = commit.evaluate(e, into_fr_plus_1(i)); return p;
    }
    // std::map iterates in key order like the BTreeMap of the reference's examples; only the first t+1 are used
    Signature combine_signatures(const Engine &e, const std::map<uint64_t, SignatureShare> &shares) const { // src/lib.rs:608-615
        size_t t = threshold();
        if (shares.size() <= t) throw NotEnoughShares();                                                    // src/lib.rs:731-733
        std::vector<uint8_t> xs, pts;
        size_t k = 0;
        for (auto &kv : shares) {
            if (k++ > t) break;
            FrBytes x = into_fr_plus_1(kv.first);
            xs.insert(xs.end(), x.begin(), x.end());
            pts.insert(pts.end(), kv.second.raw.begin(), kv.second.raw.end());
        }
        Signature out; uint8_t st = 0;
        e.check(tcb_combine_g2_batch(e.ctx(), 1, t, xs.data(), pts.data(), out.raw.data(), &st));
        if (st == 2) throw DuplicateEntry();
        if (st) throw Error("invalid encoding");
        return out;
    }
    Bytes decrypt(const Engine &e, const std::map<uint64_t, DecryptionShare> &shares, const Ciphertext &ct) const {   // src/lib.rs:618-626
        size_t t = threshold();
        if (shares.size() <= t) throw NotEnoughShares();
        std::vector<uint8_t> xs, pts;
        size_t k = 0;
        for (auto &kv : shares) {
            if (k++ > t) break;
            FrBytes x = into_fr_plus_1(kv.first);
            xs.insert(xs.end(), x.begin(), x.end());
            pts.insert(pts.end(), kv.second.raw.begin(), kv.second.raw.end());
        }
        Msgs m; m.push(ct.v);
        Bytes out(ct.v.size() ? ct.v.size() : 1);
        uint8_t st = 0;
        e.check(tcb_decrypt_batch(e.ctx(), 1, t, xs.data(), pts.data(), m.data(), m.off.data(), out.data(), &st));
        if (st == 2) throw DuplicateEntry();
        if (st) throw Error("invalid encoding");
        out.resize(ct.v.size());
        return out;
    }
};

class SecretKeySet {                                                                                        // src/lib.rs:630-688
  public:
    Poly poly;
    size_t threshold() const { return poly.degree(); }
    SecretKeyShare secret_key_share(uint64_t i) const {                                                     // src/lib.rs:669-673
        SecretKeyShare s; s.fr = poly.evaluate(Fr::from_bytes(into_fr_plus_1(i))).bytes(); return s;
    }
    PublicKeySet public_keys(const Engine &e) const {                                                       // Poly::commitment, src/poly.rs:372-377
        std::vector<uint8_t> sk;
        for (auto &c : poly.coeff) { FrBytes b = c.bytes(); sk.insert(sk.end(), b.begin(), b.end()); }
        PublicKeySet ps; ps.commit.coeff.resize(poly.coeff.size());
        e.check(tcb_g1_mul_gen_batch(e.ctx(), poly.coeff.size(), sk.data(), ps.commit.coeff[0].data()));
        return ps;
    }
};

// ---- batched entry point a throughput-minded caller uses: many (pk, sig, msg) triples, one GPU call
inline std::vector<bool> verify_batch(const Engine &e, const std::vector<PublicKey> &pks, const std::vector<Signature> &sigs, const std::vector<Bytes> &msgs) {
    size_t n = pks.size();
    std::vector<uint8_t> pk(n * 96), sg(n * 192), ok(n ? n : 1);
    Msgs m;
    for (size_t i = 0; i < n; i++) { std::memcpy(&pk[96 * i], pks[i].raw.data(), 96); std::memcpy(&sg[192 * i], sigs[i].raw.data(), 192); m.push(msgs[i]); }
    e.check(tcb_verify_batch(e.ctx(), n, pk.data(), sg.data(), m.data(), m.off.data(), ok.data()));
    std::vector<bool> out(n);
    for (size_t i = 0; i < n; i++) out[i] = ok[i] != 0;
    return out;
}

}  // namespace tcb200
