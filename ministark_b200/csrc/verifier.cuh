// verifier.cuh -- Stark::verify (src/starks.rs:171-235) + Fri::verify (src/fri.rs:191-281) on the canonical proof dump.
// The caller supplies what the reference's caller supplies -- the Constrains object (src/air.rs:163-186), here as the C
// coefficient columns on the device.  The heavy part (re-evaluating every constraint polynomial and their mix at the Q query
// points, src/starks.rs:204-224) runs on the device with the prover's own evaluation kernel; the transcript replay, the
// per-query FRI consistency checks and the Merkle paths are a few hundred field operations / hashes on the host.
// `strict` additionally enforces what the reference computes and discards (check_proof's result, fri.rs:237,239 -- against the
// root of the round the paths actually open).  Product code: nothing here touches oracle/.
#pragma once
#include "common.cuh"
#include "field.cuh"
#include "merkle.cuh"
#include "poly.cuh"
#include "transcript.hpp"

namespace ms {

struct ProofReader {
    const uint8_t* p;
    uint64_t len, pos = 0;
    bool bad = false;
    const uint8_t* take(uint64_t n) {
        if (bad || pos + n > len || pos + n < pos) { bad = true; return nullptr; }
        const uint8_t* q = p + pos;
        pos += n;
        return q;
    }
    uint64_t u64() {
        const uint8_t* q = take(8);
        uint64_t v = 0;
        if (q) memcpy(&v, q, 8);
        return v;
    }
    template <class F>
    Ext<F> ext() {
        Ext<F> e = ext_zero<F>();
        const uint8_t* q = take(sizeof(typename F::T) * F::D);
        if (q) {
            memcpy(e.c, q, sizeof(typename F::T) * F::D);
            for (int d = 0; d < F::D; d++)
                if ((uint64_t)e.c[d] >= (uint64_t)F::P) bad = true;  // ark deserialisation rejects non-canonical scalars
        }
        return e;
    }
};

// Display of an extension element (merkle.rs:165) on the host, for the leaf hash of a Merkle path
inline void host_fmt_base(uint64_t v, int zero_empty, std::string* s) {
    if (v == 0) { if (!zero_empty) s->push_back('0'); return; }
    char t[24];
    int n = 0;
    while (v) { t[n++] = (char)('0' + v % 10); v /= 10; }
    while (n) s->push_back(t[--n]);
}
template <class T>
void host_fmt_quad(const T* c, int deg, int zero_empty, std::string* s) {
    if (deg == 1) return host_fmt_base((uint64_t)c[0], zero_empty, s);
    *s += "QuadExtField(";
    host_fmt_quad(c, deg / 2, zero_empty, s);
    *s += " + ";
    host_fmt_quad(c + deg / 2, deg / 2, zero_empty, s);
    *s += " * u)";
}

// one Merkle path of the dump (merkle.rs:293-298) against `root` (MerkleRoot::check_proof, merkle.rs:312-338);
// *has_y: y is among the leaf neighbours (fri.rs:236,238)
template <class F>
bool read_and_check_path(ProofReader& rd, const Ext<F>& y, const uint8_t* root, int zero_empty, bool* has_y, bool* path_ok) {
    const uint64_t nn = rd.u64();
    if (rd.bad || nn > 1024) return false;
    Sha256Host h;
    *has_y = false;
    for (uint64_t e = 0; e < nn; e++) {
        Ext<F> v = rd.ext<F>();
        if (ext_eq(v, y)) *has_y = true;
        std::string s;
        host_fmt_quad(v.c, F::D, zero_empty, &s);
        h.update(s.data(), s.size());
    }
    uint8_t cur[32];
    h.digest(cur);
    const uint64_t levels = rd.u64();
    if (rd.bad || levels > 64) return false;
    *path_ok = true;
    for (uint64_t l = 0; l < levels; l++) {
        const uint64_t kk = rd.u64();
        if (rd.bad || kk > 64) return false;
        const uint8_t* sib = rd.take(32 * kk);
        if (!sib) return false;
        bool found = false;
        for (uint64_t s = 0; s < kk; s++) found = found || !memcmp(sib + 32 * s, cur, 32);
        if (!found) *path_ok = false;
        Sha256Host hn;
        hn.update(sib, 32 * kk);
        hn.digest(cur);
    }
    if (root && memcmp(cur, root, 32)) *path_ok = false;
    return true;
}

// *accepted = 1 / 0; *failed_line = the reference line of the first failed check (0 when accepted, -2 malformed dump)
template <class F>
int stark_verify(Ctx* c, const ms_stark_params& p, const typename F::T* d_constrains, uint64_t stride, uint64_t n, uint64_t cols,
                 const uint8_t* proof, uint64_t proof_len, int strict, int32_t* accepted, int32_t* failed_line) {
    using T = typename F::T;
    using E = Ext<F>;
    constexpr int D = F::D;
    if (!accepted || !proof) return fail(c, MS_ERR_BAD_SHAPE, "ms_stark_verify: null argument");
    *accepted = 0;
    auto reject = [&](int line) { if (failed_line) *failed_line = line; return MS_OK; };
    StarkDerived der;
    if (stark_derive(F::ID, p, &der) != MS_OK) return fail(c, MS_ERR_BAD_SHAPE, "bad STARK parameters (starks.rs:317-320)");
    const uint64_t R = der.rounds, Q = der.constrain_queries, QF = der.fri_queries;
    ProofReader rd{proof, proof_len};
    const uint8_t* magic = rd.take(8);
    if (!magic || memcmp(magic, "MSTARKP1", 8)) return reject(-2);
    const uint8_t* hdr = rd.take(8);
    uint32_t h2[2] = {0, 0};
    if (hdr) memcpy(h2, hdr, 8);
    if (!hdr || h2[0] != (uint32_t)F::ID || h2[1] != (uint32_t)D) return reject(-2);
    const uint64_t alen = rd.u64();
    const uint8_t* arthur = rd.take(alen);
    const uint8_t* trace_commit = rd.take(32);
    const uint8_t* lde_commit = rd.take(32);
    if (rd.bad) return reject(-2);
    // ---- Arthur: replay the transcript over the absorbed bytes (starks.rs:186-199)
    IOPattern io = stark_iopattern(F::BITS, D, R, Q, QF);
    Merlin ar(io, c->bridge_masks, c->leftover_as_published != 0);
    uint64_t apos = 0;
    auto next_bytes = [&](uint8_t* out, size_t k) -> bool {  // Arthur::fill_next_bytes: read, then absorb
        if (apos + k > alen) return false;
        memcpy(out, arthur + apos, k);
        apos += k;
        return ar.add_bytes(out, k);
    };
    auto chal = [&](int degree, size_t count, uint64_t* out) { return ar.challenge_scalars(F::BITS, (uint64_t)F::P, degree, count, out); };
    uint8_t d32[32];
    if (!next_bytes(d32, 32) || memcmp(d32, trace_commit, 32)) return reject(187);
    uint64_t shift64, r64;
    if (!chal(1, 1, &shift64)) return reject(189);
    if (!next_bytes(d32, 32) || memcmp(d32, lde_commit, 32)) return reject(191);
    if (!chal(1, 1, &r64)) return reject(193);
    std::vector<uint64_t> zraw(Q * D);
    if (!chal(D, Q, zraw.data())) return reject(199);
    std::vector<E> zq(Q);
    for (uint64_t q = 0; q < Q; q++)
        for (int d = 0; d < D; d++) zq[q].c[d] = (T)zraw[q * D + d];
    // ---- re-evaluate the constraint polynomials and their mix at the query points on the device (starks.rs:204-224)
    {
        const uint64_t nq_ = rd.u64(), nc_ = rd.u64();
        if (rd.bad || nq_ != Q || (Q && nc_ != cols)) return reject(209);
        uint64_t dom = 1;
        while (dom < p.steps) dom <<= 1;  // Radix2EvaluationDomain::new(degree + 1), starks.rs:190
        Scratch mixed(c), dlen(c);
        MS_TRY(mixed.alloc(n * sizeof(T)));
        MS_TRY(dlen.alloc(8));
        MS_CUDA(c, cudaMemsetAsync(dlen.p, 0, 8, c->stream));
        MS_TRY(mix<F>(c, d_constrains, stride, n, cols, (T)r64, mixed.as<T>()));
        k_poly_len<F><<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(mixed.as<T>(), n, 1, n, dlen.as<unsigned long long>());
        MS_LAUNCH_CHECK(c);
        std::vector<E> got(Q * cols), gotv(Q);
        MS_TRY(eval_points<F>(c, d_constrains, stride, 0, 1, n, 1, cols, zq.data(), (int)Q, got.data()));
        MS_TRY(eval_points<F>(c, mixed.as<T>(), n, 0, 1, n, 1, 1, zq.data(), (int)Q, gotv.data()));
        unsigned long long cxl = 0;
        MS_CUDA(c, cudaMemcpyAsync(&cxl, dlen.p, 8, cudaMemcpyDeviceToHost, c->stream));
        MS_CUDA(c, cudaStreamSynchronize(c->stream));
        if (cxl > dom) return reject(220);  // the "quotient" by the vanishing polynomial must be zero (starks.rs:220-221)
        for (uint64_t q = 0; q < Q; q++)
            for (uint64_t col = 0; col < cols; col++) {
                E claimed = rd.ext<F>();
                if (rd.bad || !ext_eq(claimed, got[q * cols + col])) return reject(216);
            }
        if (rd.u64() != Q) return reject(223);
        for (uint64_t q = 0; q < Q; q++) {
            E claimed = rd.ext<F>();
            if (rd.bad || !ext_eq(claimed, gotv[q])) return reject(223);
        }
    }
    // ---- Fri::verify (fri.rs:191-281)
    std::vector<E> alphas(R), deepq(R), deepp(2 * R);
    std::vector<uint8_t> commits(32 * R);
    for (uint64_t i = 0; i + 1 < R; i++) {  // fri.rs:258-270
        uint64_t raw[4];
        if (!chal(D, 1, raw)) return reject(260);
        for (int d = 0; d < D; d++) deepq[i].c[d] = (T)raw[d];
        uint8_t db[2 * 4 * 8];
        if (!next_bytes(db, 2 * sizeof(E))) return reject(262);
        memcpy(&deepp[2 * i], db, 2 * sizeof(E));
        if (!chal(D, 1, raw)) return reject(266);
        for (int d = 0; d < D; d++) alphas[i].c[d] = (T)raw[d];
        if (!next_bytes(&commits[32 * i], 32)) return reject(268);
    }
    std::vector<uint8_t> braw(8 * QF, 0);
    if (!ar.challenge_bytes(braw.data(), braw.size())) return reject(273);
    std::vector<uint64_t> betas(QF);
    const uint64_t domain_size = 1ULL << R;
    for (uint64_t k = 0; k < QF; k++) {
        memcpy(&betas[k], &braw[8 * k], 8);
        if (betas[k] > domain_size) betas[k] %= domain_size;  // fri.rs:277
    }
    const T g0 = root_of_unity<F>((int)R);  // fri.rs:209
    std::vector<E> prev_x3(QF);
    for (uint64_t k = 0; k < QF; k++) prev_x3[k] = ext_from_base<F>(fpow<F>(g0, betas[k]));  // fri.rs:210
    if (rd.u64() != R - 1) return reject(206);
    for (uint64_t i = 0; i + 1 < R; i++) {
        if (rd.u64() != QF) return reject(207);
        for (uint64_t k = 0; k < QF; k++) {
            E pt[6];
            for (int e = 0; e < 6; e++) pt[e] = rd.ext<F>();
            if (rd.bad) return reject(-2);
            const E x1 = pt[0], y1 = pt[1], x2 = pt[2], y2 = pt[3], x3 = pt[4], y3 = pt[5];
            if (!ext_eq(x1, prev_x3[k])) return reject(217);
            if (!ext_eq(ext_sub(ext_zero<F>(), x1), x2)) return reject(218);
            if (!ext_eq(ext_mul(x1, x1), x3)) return reject(219);
            for (int which = 0; which < 2; which++) {
                bool has_y = false, path_ok = false;
                // the paths open round i, whose root is commits[i-1] (round 0's root is not in the transcript, fri.rs:77-82)
                const uint8_t* root = (strict && i > 0) ? &commits[32 * (i - 1)] : nullptr;
                if (!read_and_check_path<F>(rd, which ? y2 : y1, root, c->zero_display_empty, &has_y, &path_ok)) return reject(-2);
                if (!has_y) return reject(which ? 238 : 236);
                if (strict && i > 0 && !path_ok) return reject(237);
            }
            const uint64_t nq = rd.u64();
            if (rd.bad || nq > (rd.len - rd.pos) / sizeof(E)) return reject(-2);  // a tampered count: nq * sizeof(E) must not wrap
            const uint8_t* qraw = rd.take(nq * sizeof(E));
            if (rd.bad) return reject(-2);
            uint64_t qlen = nq;  // trailing zero coefficients do not count (DensePolynomial)
            while (qlen > 0) {
                bool z = true;
                for (size_t b = 0; b < sizeof(E); b++) z = z && !qraw[(qlen - 1) * sizeof(E) + b];
                if (!z) break;
                qlen--;
            }
            const uint64_t q_deg = qlen ? qlen - 1 : 0, total_degree = q_deg + 3;  // fri.rs:223-224
            if (total_degree < 2 || total_degree > (1ULL << (R - i))) return reject(225);
            const T dinv = finv<F>(F::sub(x2.c[0], x1.c[0]));
            const E a = ext_mul_base(ext_sub(y2, y1), dinv);                             // fri.rs:229
            const E b = ext_sub(y1, ext_mul(a, x1));                                     // fri.rs:230
            const E deep_at_alpha = ext_add(deepp[2 * i], ext_mul(deepp[2 * i + 1], alphas[i]));
            const E deep_adj = ext_add(ext_mul(y3, ext_sub(x3, deepq[i])), deep_at_alpha);   // fri.rs:231-232
            if (!ext_eq(ext_add(b, ext_mul(a, alphas[i])), deep_adj)) return reject(233);
            prev_x3[k] = x3;                                                             // fri.rs:240
        }
    }
    if (rd.bad || rd.pos != rd.len) return reject(-2);
    *accepted = 1;
    if (failed_line) *failed_line = 0;
    return MS_OK;
}

}  // namespace ms
