// ministark.hpp -- C++17 host mirror of the reference crate's public API for the prover path, over the C ABI of
// ministark.h (header only; link libministark.so).
//
// The reference is compiled code (Rust) whose toolchain is absent from this image, so the host side above the C ABI is given
// in C++ as well as in Python (ministark_b200/air.py, starks.py): same names, argument meaning and error behaviour as
//   src/air.rs     Provable<W>, TraceTable (new / step_number / add_row / add_boundary_constrain / add_transition_constrain /
//                  constrain_number / derive_constrains), Constrains (len / is_empty / get_constrain_poly)
//   src/starks.rs  StarkConfig::new(security_bits, blowup_factor, steps, trace_columns), Stark::new / prove / verify, StarkProof
//   src/field.rs   Goldilocks, BabyBear (constants only: the arithmetic of the path runs on the device)
// so that tests/e2e_goldilocks.rs and tests/e2e_babybear.rs read the same here (examples/e2e_fibonacci.cpp).  Transition
// constraints stay host closures over DensePolynomial (src/air.rs:61); only constraints affine in the trace polynomials are
// provable by the reference (src/starks.rs:119), and TraceTable::affine_form() recovers their T x W scalar matrix and constants by
// probing the closures -- that matrix is what crosses the boundary (ms_stark_prove_affine).  A reference `assert!` / panic is a
// std::logic_error here, an Err(ProverError / VerifierError) a ministark::Error carrying the MS_* code.  There is no CPU fallback:
// without a CUDA device Gpu's constructor throws.
#pragma once
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "ministark.h"

namespace ministark {

using u64 = uint64_t;
using u128 = unsigned __int128;

struct Error : std::runtime_error {
    int32_t code;
    Error(int32_t c, const std::string& what) : std::runtime_error(what), code(c) {}
};
#define MINISTARK_ASSERT(cond, msg)                                         \
    do {                                                                    \
        if (!(cond)) throw std::logic_error(std::string("assertion failed: ") + (msg)); \
    } while (0)

// ------------------------------------------------------------------------------------------------ src/field.rs:43-109
struct StarkField {
    const char* name;
    int32_t id;        // MS_FIELD_*
    u64 p;
    u64 generator;     // multiplicative generator; the two-adic root is generator^((p-1) >> two_adicity)
    int two_adicity;
    int ext_degree;    // F::Extension: Fp2 (Goldilocks), Fp4 (BabyBear)

    int modulus_bits() const { int b = 0; for (u64 v = p; v; v >>= 1) b++; return b; }
    size_t base_bytes() const { return (size_t)(modulus_bits() + 7) / 8; }   // ark-serialize compressed size
    size_t elem_bytes() const { return id == MS_FIELD_GOLDILOCKS ? 8 : 4; }  // element size across the C ABI
    u64 add(u64 a, u64 b) const { return (u64)(((u128)a + b) % p); }
    u64 sub(u64 a, u64 b) const { return (u64)(((u128)a + p - b % p) % p); }
    u64 mul(u64 a, u64 b) const { return (u64)((u128)a * b % p); }
    u64 pow(u64 b, u64 e) const {
        u64 r = 1 % p;
        for (b %= p; e; e >>= 1, b = mul(b, b))
            if (e & 1) r = mul(r, b);
        return r;
    }
    u64 inv(u64 a) const { return pow(a, p - 2); }
    // ark-poly Radix2EvaluationDomain::group_gen for size 2^log_n
    u64 root_of_unity(int log_n) const {
        MINISTARK_ASSERT(log_n >= 0 && log_n <= two_adicity, "domain exceeds the field's two-adicity");
        return pow(pow(generator, (p - 1) >> two_adicity), 1ULL << (two_adicity - log_n));
    }
};
inline const StarkField& Goldilocks() {
    static const StarkField f{"Goldilocks", MS_FIELD_GOLDILOCKS, 0xFFFFFFFF00000001ULL, 7, 32, 2};
    return f;
}
inline const StarkField& BabyBear() {
    static const StarkField f{"BabyBear", MS_FIELD_BABYBEAR, 2013265921ULL, 440564289ULL, 27, 4};
    return f;
}

// ------------------------------------------------------------------------------------------------ ark-poly DensePolynomial
// Just enough for constraint closures: +, -, * and scalar products; coefficients trimmed of trailing zeros.
class DensePolynomial {
  public:
    DensePolynomial(const StarkField& F, std::vector<u64> c) : F_(&F), coeffs(std::move(c)) {
        for (auto& v : coeffs) v %= F.p;
        while (!coeffs.empty() && coeffs.back() == 0) coeffs.pop_back();
    }
    static DensePolynomial from_coefficients_vec(const StarkField& F, std::vector<u64> c) { return DensePolynomial(F, std::move(c)); }
    DensePolynomial clone() const { return *this; }
    size_t degree() const { return coeffs.empty() ? 0 : coeffs.size() - 1; }
    bool is_zero() const { return coeffs.empty(); }
    const StarkField& field() const { return *F_; }
    DensePolynomial operator+(const DensePolynomial& o) const { return zip(o, false); }
    DensePolynomial operator-(const DensePolynomial& o) const { return zip(o, true); }
    DensePolynomial operator*(const DensePolynomial& o) const {
        std::vector<u64> out(coeffs.empty() || o.coeffs.empty() ? 0 : coeffs.size() + o.coeffs.size() - 1, 0);
        for (size_t i = 0; i < coeffs.size(); i++)
            for (size_t j = 0; j < o.coeffs.size(); j++) out[i + j] = F_->add(out[i + j], F_->mul(coeffs[i], o.coeffs[j]));
        return DensePolynomial(*F_, std::move(out));
    }
    DensePolynomial operator*(u64 s) const {
        std::vector<u64> out(coeffs);
        for (auto& v : out) v = F_->mul(v, s % F_->p);
        return DensePolynomial(*F_, std::move(out));
    }
    bool operator==(const DensePolynomial& o) const { return coeffs == o.coeffs; }
    bool operator!=(const DensePolynomial& o) const { return !(*this == o); }

  private:
    const StarkField* F_;
    DensePolynomial zip(const DensePolynomial& o, bool minus) const {
        std::vector<u64> out(std::max(coeffs.size(), o.coeffs.size()), 0);
        for (size_t i = 0; i < out.size(); i++) {
            const u64 a = i < coeffs.size() ? coeffs[i] : 0, b = i < o.coeffs.size() ? o.coeffs[i] : 0;
            out[i] = minus ? F_->sub(a, b) : F_->add(a, b);
        }
        return DensePolynomial(*F_, std::move(out));
    }

  public:
    std::vector<u64> coeffs;
};

// ------------------------------------------------------------------------------------------------ ark_std::test_rng
// `F::rand(&mut test_rng())` (src/air.rs:81): ark-std's fixed-seed StdRng (rand 0.8: ChaCha12), one fresh generator per
// cell, so every padding cell holds this same value.  ark-ff's Fp::rand masks the draw to the modulus bit length, rejects
// values >= p and takes the draw as the Montgomery representation (R = 2^64).  Recalled from upstream,
// unpinned; only an INPUT of the device path.  Twin of ministark_b200/air.py:padding_value.
inline void chacha_block(const uint32_t key[8], u64 counter, int rounds, uint32_t out[16]) {
    auto rotl = [](uint32_t x, int n) { return (x << n) | (x >> (32 - n)); };
    uint32_t st[16] = {0x61707865u, 0x3320646Eu, 0x79622D32u, 0x6B206574u};
    for (int i = 0; i < 8; i++) st[4 + i] = key[i];
    st[12] = (uint32_t)counter; st[13] = (uint32_t)(counter >> 32); st[14] = 0; st[15] = 0;
    uint32_t x[16];
    std::memcpy(x, st, sizeof x);
    auto qr = [&](int a, int b, int c, int d) {
        x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 16);
        x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 12);
        x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 8);
        x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 7);
    };
    for (int i = 0; i < rounds / 2; i++) {
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15);
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14);
    }
    for (int i = 0; i < 16; i++) out[i] = x[i] + st[i];
}
inline u64 padding_value(const StarkField& F) {
    const uint8_t seed[32] = {1, 0, 0, 0, 23, 0, 0, 0, 200, 1, 0, 0, 210, 30, 0, 0};
    uint32_t key[8];
    for (int i = 0; i < 8; i++) key[i] = (uint32_t)seed[4 * i] | (uint32_t)seed[4 * i + 1] << 8 | (uint32_t)seed[4 * i + 2] << 16 | (uint32_t)seed[4 * i + 3] << 24;
    const int bits = F.modulus_bits();
    const u64 mask = bits >= 64 ? ~0ULL : ((1ULL << bits) - 1);
    const u64 rinv = F.inv((u64)(((u128)1 << 64) % F.p));
    for (u64 counter = 0;; counter++) {
        uint32_t blk[16];
        chacha_block(key, counter, 12, blk);
        for (int i = 0; i + 1 < 16; i += 2) {  // a block holds 8 whole 64-bit draws: none straddles two blocks
            const u64 v = (((u64)blk[i + 1] << 32) | blk[i]) & mask;
            if (v < F.p) return F.mul(v, rinv);
        }
    }
}

// ------------------------------------------------------------------------------------------------ device handle
// One context (GPU, stream, field) of the library; RAII over ms_ctx_create / ms_ctx_destroy.
class Gpu {
  public:
    explicit Gpu(const StarkField& F, int device = 0) : F_(&F) {
        const int32_t rc = ms_ctx_create(F.id, device, nullptr, &ctx_);
        if (rc != MS_OK || !ctx_) throw Error(rc, "ms_ctx_create failed: no usable CUDA device (there is no CPU fallback)");
    }
    ~Gpu() { if (ctx_) ms_ctx_destroy(ctx_); }
    Gpu(const Gpu&) = delete;
    Gpu& operator=(const Gpu&) = delete;
    ms_ctx* ctx() const { return ctx_; }
    const StarkField& field() const { return *F_; }
    void check(int32_t rc, const char* what) const {
        if (rc == MS_OK) return;
        const char* e = ms_last_error(ctx_);
        const std::string msg = std::string(what) + ": " + (e && *e ? e : "error") + " (code " + std::to_string(rc) + ")";
        // shape violations and the non-zero remainder are panics in the reference (src/merkle.rs:95,99-104, src/air.rs:23, src/starks.rs:84-85,119)
        if (rc == MS_ERR_BAD_SHAPE || rc == MS_ERR_QUOTIENT_NONZERO) throw std::logic_error(msg);
        throw Error(rc, msg);
    }
    // host vector of canonical u64 values <-> the element type of the ABI (uint64 Goldilocks, uint32 BabyBear)
    std::vector<uint8_t> pack(const std::vector<u64>& v) const {
        std::vector<uint8_t> out(v.size() * F_->elem_bytes());
        if (F_->elem_bytes() == 8) std::memcpy(out.data(), v.data(), out.size());
        else for (size_t i = 0; i < v.size(); i++) { const uint32_t x = (uint32_t)v[i]; std::memcpy(&out[4 * i], &x, 4); }
        return out;
    }
    std::vector<u64> unpack(const std::vector<uint8_t>& raw) const {
        std::vector<u64> out(raw.size() / F_->elem_bytes());
        if (F_->elem_bytes() == 8) std::memcpy(out.data(), raw.data(), raw.size());
        else for (size_t i = 0; i < out.size(); i++) { uint32_t x; std::memcpy(&x, &raw[4 * i], 4); out[i] = x; }
        return out;
    }

  private:
    const StarkField* F_;
    ms_ctx* ctx_ = nullptr;
};
// device memory owned through the ABI (ms_dev_alloc / ms_dev_free)
class DeviceBuffer {
  public:
    DeviceBuffer(const Gpu& g, size_t bytes) : g_(&g), bytes_(bytes) { g.check(ms_dev_alloc(g.ctx(), bytes ? bytes : 16, &p_), "ms_dev_alloc"); }
    ~DeviceBuffer() { if (p_) ms_dev_free(g_->ctx(), p_); }
    DeviceBuffer(const DeviceBuffer&) = delete;
    DeviceBuffer& operator=(const DeviceBuffer&) = delete;
    void* ptr() const { return p_; }
    size_t bytes() const { return bytes_; }

  private:
    const Gpu* g_;
    void* p_ = nullptr;
    size_t bytes_;
};

// ------------------------------------------------------------------------------------------------ src/air.rs:163-186
// The constraint polynomials (trace polynomials first) as `len()` coefficient columns of length n on the device.
class Constrains {
  public:
    Constrains(std::shared_ptr<Gpu> gpu, size_t trace_n, size_t transition_n, u64 n, std::shared_ptr<DeviceBuffer> polys)
        : gpu_(std::move(gpu)), trace_constrains_num(trace_n), transition_constrains_num(transition_n), n_(n), polys_(std::move(polys)) {}
    size_t len() const { return trace_constrains_num + transition_constrains_num; }   // air.rs:169-171
    bool is_empty() const { return len() == 0; }                                      // air.rs:173-175
    std::vector<u64> get_constrain_poly(size_t index) const {                        // air.rs:178-181 (trimmed like DensePolynomial)
        MINISTARK_ASSERT(index < len(), "constraint index out of range");
        const size_t eb = gpu_->field().elem_bytes();
        std::vector<uint8_t> raw(n_ * eb);
        gpu_->check(ms_d2h(gpu_->ctx(), raw.data(), static_cast<const uint8_t*>(polys_->ptr()) + index * n_ * eb, raw.size()), "ms_d2h");
        std::vector<u64> c = gpu_->unpack(raw);
        while (!c.empty() && c.back() == 0) c.pop_back();
        return c;
    }
    const void* device_polynomials() const { return polys_->ptr(); }                  // air.rs:183-185 get_polynomials
    u64 length() const { return n_; }
    const std::shared_ptr<Gpu>& gpu() const { return gpu_; }

  private:
    std::shared_ptr<Gpu> gpu_;

  public:
    size_t trace_constrains_num, transition_constrains_num;

  private:
    u64 n_;
    std::shared_ptr<DeviceBuffer> polys_;
};

// ------------------------------------------------------------------------------------------------ src/air.rs:61-161
using Constrain = std::function<DensePolynomial(const std::vector<DensePolynomial>&)>;

class TraceTable {
  public:
    TraceTable(const StarkField& F, size_t steps, size_t registers) : F_(&F), steps_(steps), width_(registers) {
        const size_t n = steps + 1;  // Radix2EvaluationDomain::new(steps + 1), air.rs:74
        length_ = 1;
        int log_n = 0;
        while (length_ < n) { length_ <<= 1; log_n++; }
        omega = F.root_of_unity(log_n);
        data_.assign(length_ * registers, 0);                       // row-major Matrix (air.rs:15-59)
        const u64 pad = padding_value(F);                           // air.rs:77-83
        for (size_t i = steps * registers; i < data_.size(); i++) data_[i] = pad;
    }
    static TraceTable new_(const StarkField& F, size_t steps, size_t registers) { return TraceTable(F, steps, registers); }
    size_t step_number() const { return steps_; }
    size_t length() const { return length_; }
    size_t width() const { return width_; }
    const StarkField& field() const { return *F_; }
    const std::vector<u64>& data() const { return data_; }
    void add_row(size_t index, const std::vector<u64>& row) {      // air.rs:106-112
        MINISTARK_ASSERT(row.size() == width_, "row.len() == trace.width");
        MINISTARK_ASSERT(index < steps_, "index < steps");
        for (size_t j = 0; j < width_; j++) data_[index * width_ + j] = row[j] % F_->p;
    }
    void add_boundary_constrain(size_t row, size_t col) {           // air.rs:114-117 (recorded, never used by the reference)
        MINISTARK_ASSERT(row < steps_ && col < width_, "row < steps && col < trace.width");
        boundaries_.emplace_back(row, col);
    }
    void add_transition_constrain(Constrain f) { transition_constrains_.push_back(std::move(f)); }  // air.rs:119-121
    size_t constrain_number() const { return width_ + transition_constrains_.size(); }             // air.rs:123-125

    // (M, c): closure t equals sum_w M[t][w] * trace_poly_w + c[t].  Throws std::logic_error if a closure is not affine in the
    // trace polynomials (a product of two of them, or a multiplication by a non-constant polynomial, reaches degree >= N and
    // makes the reference panic at starks.rs:119).
    void affine_form(std::vector<u64>* matrix, std::vector<u64>* constants) const {
        const StarkField& F = *F_;
        const size_t W = width_, T = transition_constrains_.size();
        const DensePolynomial zero(F, {}), one(F, {1});
        std::vector<DensePolynomial> zeros(W, zero), probe;
        for (size_t j = 0; j < W; j++) probe.emplace_back(F, std::vector<u64>{3 + 5 * j, 7 + j, 11 * (j + 1)});
        matrix->assign(T * W, 0);
        constants->assign(T, 0);
        for (size_t t = 0; t < T; t++) {
            const Constrain& f = transition_constrains_[t];
            const DensePolynomial c0 = f(zeros);
            MINISTARK_ASSERT(c0.coeffs.size() <= 1, "transition constraint adds a non-constant polynomial: not expressible across the C ABI");
            (*constants)[t] = c0.coeffs.empty() ? 0 : c0.coeffs[0];
            for (size_t j = 0; j < W; j++) {
                std::vector<DensePolynomial> unit;
                unit.reserve(W);
                for (size_t k = 0; k < W; k++) unit.push_back(k == j ? one : zero);
                const DensePolynomial r = f(unit) - c0;
                MINISTARK_ASSERT(r.coeffs.size() <= 1, "transition constraint multiplies by a non-constant polynomial");
                (*matrix)[t * W + j] = r.coeffs.empty() ? 0 : r.coeffs[0];
            }
            DensePolynomial want = c0;
            for (size_t j = 0; j < W; j++) want = want + probe[j] * (*matrix)[t * W + j];
            MINISTARK_ASSERT(f(probe) == want, "transition constraint is not affine in the trace polynomials");
        }
    }

    // air.rs:127-144: the trace polynomials (iNTT of every column, air.rs:147-160) followed by the transition polynomials,
    // computed on the device (ms_intt_columns + ms_linear_constraints) and kept there.
    Constrains derive_constrains(std::shared_ptr<Gpu> gpu) const {
        const StarkField& F = *F_;
        MINISTARK_ASSERT(gpu->field().id == F.id, "context of another field");
        const size_t eb = F.elem_bytes(), W = width_, T = transition_constrains_.size(), n = length_;
        std::vector<u64> cm(W * n);
        for (size_t r = 0; r < n; r++)
            for (size_t c = 0; c < W; c++) cm[c * n + r] = data_[r * W + c];
        DeviceBuffer evals(*gpu, W * n * eb);
        auto polys = std::make_shared<DeviceBuffer>(*gpu, (W + T) * n * eb);
        const std::vector<uint8_t> packed = gpu->pack(cm);
        gpu->check(ms_h2d(gpu->ctx(), evals.ptr(), packed.data(), packed.size()), "ms_h2d");
        gpu->check(ms_intt_columns(gpu->ctx(), evals.ptr(), n, n, W, polys->ptr(), n), "ms_intt_columns");
        if (T) {
            std::vector<u64> m, consts;
            affine_form(&m, &consts);
            uint8_t* d_cons = static_cast<uint8_t*>(polys->ptr()) + W * n * eb;
            const std::vector<uint8_t> pm = gpu->pack(m);
            gpu->check(ms_linear_constraints(gpu->ctx(), polys->ptr(), n, n, W, pm.data(), T, d_cons, n), "ms_linear_constraints");
            for (size_t t = 0; t < T; t++) {  // + the constant polynomial: coefficient 0 of column t
                if (!consts[t]) continue;
                std::vector<uint8_t> one(eb);
                gpu->check(ms_d2h(gpu->ctx(), one.data(), d_cons + t * n * eb, eb), "ms_d2h");
                const std::vector<uint8_t> upd = gpu->pack({F.add(gpu->unpack(one)[0], consts[t])});
                gpu->check(ms_h2d(gpu->ctx(), d_cons + t * n * eb, upd.data(), eb), "ms_h2d");
            }
        }
        gpu->check(ms_sync(gpu->ctx()), "ms_sync");
        return Constrains(std::move(gpu), W, T, n, std::move(polys));
    }

    u64 omega;

  private:
    const StarkField* F_;
    size_t steps_, width_, length_ = 1;
    std::vector<u64> data_;
    std::vector<std::pair<size_t, size_t>> boundaries_;
    std::vector<Constrain> transition_constrains_;
};

// src/air.rs:9-12: `fn trace(&self, witness: &W) -> TraceTable<F>`
template <class W>
struct Provable {
    virtual ~Provable() = default;
    virtual TraceTable trace(const W& witness) const = 0;
};

// ------------------------------------------------------------------------------------------------ src/starks.rs:21-28
// The proof as the canonical dump of DESIGN.md section 6 (what crosses the C ABI); the two commitments and `arthur` -- the only
// byte string the reference itself defines -- are parsed out, the rest stays in `raw` (ministark_b200/starks.py parses all of it).
struct StarkProof {
    std::vector<uint8_t> raw;
    std::vector<uint8_t> arthur;
    uint8_t trace_commit[32];
    uint8_t constrain_trace_commit[32];

    static StarkProof from_dump(std::vector<uint8_t> raw) {
        StarkProof p;
        MINISTARK_ASSERT(raw.size() >= 24 + 64 && std::memcmp(raw.data(), "MSTARKP1", 8) == 0, "not a proof dump");
        u64 alen;
        std::memcpy(&alen, raw.data() + 16, 8);
        MINISTARK_ASSERT(alen <= raw.size() - 24 - 64, "truncated proof dump");  // (not 24 + alen + 64: a tampered length must not wrap)
        p.arthur.assign(raw.begin() + 24, raw.begin() + 24 + (size_t)alen);
        std::memcpy(p.trace_commit, raw.data() + 24 + alen, 32);
        std::memcpy(p.constrain_trace_commit, raw.data() + 24 + alen + 32, 32);
        p.raw = std::move(raw);
        return p;
    }
};

// ------------------------------------------------------------------------------------------------ src/starks.rs:236-332
// StarkConfig::new(security_bits, blowup_factor, steps, trace_columns).  `inner_children` is this repo's extension: the
// reference hard-wires binary trees (starks.rs:283-302).
class StarkConfig {
  public:
    StarkConfig(const StarkField& F, size_t security_bits, size_t blowup_factor, size_t steps, size_t trace_columns, size_t inner_children = 2)
        : F_(&F), params{security_bits, blowup_factor, steps, trace_columns, inner_children} {
        u64 r = 0, cq = 0, fq = 0;
        const int32_t rc = ms_stark_derive(F.id, &params, &r, &cq, &fq);
        if (rc != MS_OK) throw std::logic_error("StarkConfig: security bits has to be at least 20 (starks.rs:318) / bad parameters");
        rounds = r; constrain_queries = cq; fri_queries = fq;
        degree = steps - 1;
    }
    static StarkConfig new_(const StarkField& F, size_t security_bits, size_t blowup_factor, size_t steps, size_t trace_columns) {
        return StarkConfig(F, security_bits, blowup_factor, steps, trace_columns);
    }
    const StarkField& field() const { return *F_; }

  private:
    const StarkField* F_;

  public:
    ms_stark_params params;
    u64 rounds = 0, constrain_queries = 0, fri_queries = 0, degree = 0;
};

// ------------------------------------------------------------------------------------------------ src/starks.rs:30-235
class Stark {
  public:
    explicit Stark(StarkConfig config, std::shared_ptr<Gpu> gpu = nullptr)
        : config_(std::move(config)), gpu_(gpu ? std::move(gpu) : std::make_shared<Gpu>(config_.field())) {}
    static Stark new_(StarkConfig config) { return Stark(std::move(config)); }
    const std::shared_ptr<Gpu>& gpu() const { return gpu_; }
    const StarkConfig& config() const { return config_; }

    // Stark::prove (starks.rs:59-169): everything behind `air.trace(&witness)` runs on the GPU (plus the serial host
    // transcript inside the library).  Err(ProverError) -> ministark::Error, a reference panic -> std::logic_error.
    template <class W>
    StarkProof prove(const Provable<W>& air, const W& witness) const {
        const TraceTable trace = air.trace(witness);
        MINISTARK_ASSERT(trace.field().id == config_.field().id, "trace over another field");
        std::vector<u64> m, consts;
        trace.affine_form(&m, &consts);
        const size_t T = consts.size();
        bool any_const = false;
        for (u64 c : consts) any_const = any_const || c != 0;
        const std::vector<uint8_t> tr = gpu_->pack(trace.data()), pm = gpu_->pack(m), pc = gpu_->pack(consts);
        u64 cap = ms_stark_proof_bound(config_.field().id, &config_.params, trace.length(), trace.constrain_number());
        MINISTARK_ASSERT(cap > 0, "StarkConfig does not fit the trace");
        std::vector<uint8_t> buf(cap);
        u64 len = cap;
        gpu_->check(ms_stark_prove_affine(gpu_->ctx(), &config_.params, tr.data(), trace.length(), trace.width(), pm.data(),
                                          any_const ? pc.data() : nullptr, T, buf.data(), &len), "Stark::prove");
        buf.resize(len);
        return StarkProof::from_dump(std::move(buf));
    }

    // Stark::verify (starks.rs:171-235).  Ok(true) -> true; a failed `assert!` of the reference -> std::logic_error naming the
    // reference line; a malformed dump -> ministark::Error.  strict also enforces the Merkle paths (fri.rs:237,239 discards them).
    bool verify(const Constrains& constrains, const StarkProof& proof, bool strict = false) const {
        int32_t accepted = 0, line = 0;
        gpu_->check(ms_stark_verify(gpu_->ctx(), &config_.params, constrains.device_polynomials(), constrains.length(), constrains.length(),
                                    constrains.len(), proof.raw.data(), proof.raw.size(), strict ? 1 : 0, &accepted, &line), "Stark::verify");
        if (accepted) return true;
        if (line > 0) throw std::logic_error("Stark::verify: check at reference line " + std::to_string(line) + " failed");
        throw Error(MS_ERR_TRANSCRIPT, "Stark::verify: malformed proof dump");
    }

  private:
    StarkConfig config_;
    std::shared_ptr<Gpu> gpu_;
};

}  // namespace ministark
