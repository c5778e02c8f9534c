//! `crate::gpu` -- safe wrapper of libministark.so for the reference crate (feature `b200`).
//!
//! What crosses the boundary (include/ministark.h): the padded row-major trace `air.trace(&witness)` built (canonical
//! integers), the T x W scalar matrix the AIR's transition closures denote, and the `StarkConfig::new` arguments; what comes
//! back is the canonical proof dump (DESIGN.md "Proof bytes"), parsed into the crate's own `StarkProof`.  Everything
//! between "trace exists" (src/starks.rs:68) and "proof returned" (src/starks.rs:161-168) runs on the GPU(s).
//!
//! NOT COMPILED in this repository's image (no cargo / rustc); written against ark-ff / ark-poly 0.5.0.
pub mod ffi;
pub mod starks_gpu;

use ark_ff::{BigInteger, PrimeField};
use ark_poly::univariate::DensePolynomial;
use ark_poly::DenseUVPolynomial;
use std::ffi::CStr;
use std::ptr;

use crate::air::TraceTable;
use crate::error::ProverError;

/// Which of the two fields of src/field.rs:43-109 a `StarkField` marker stands for on the device.
pub trait GpuField: crate::field::StarkField {
    const FIELD_ID: i32;
}
impl GpuField for crate::field::Goldilocks {
    const FIELD_ID: i32 = ffi::MS_FIELD_GOLDILOCKS;
}
impl GpuField for crate::field::BabyBear {
    const FIELD_ID: i32 = ffi::MS_FIELD_BABYBEAR;
}

/// One prover context = one GPU + one stream (`ms_ctx`).  Not `Sync`: one prover per context, like the C ABI says.
pub struct Gpu {
    pub(crate) ctx: *mut ffi::ms_ctx,
    pub(crate) field: i32,
}
unsafe impl Send for Gpu {}

impl Gpu {
    pub fn new(field: i32, device: i32) -> Result<Self, i32> {
        let mut ctx = ptr::null_mut();
        let rc = unsafe { ffi::ms_ctx_create(field, device, ptr::null_mut(), &mut ctx) };
        if rc == ffi::MS_OK { Ok(Self { ctx, field }) } else { Err(rc) }
    }
    pub fn last_error(&self) -> String {
        unsafe { CStr::from_ptr(ffi::ms_last_error(self.ctx)) }.to_string_lossy().into_owned()
    }
    /// Join a multi-GPU group, one process per GPU: `id` comes from `Gpu::unique_id()` on rank 0, sent to the other ranks by
    /// any means; afterwards every `prove` on these contexts is one sharded proof (all ranks must call it).
    pub fn join_nccl(&mut self, id: &[u8; 128], rank: i32, world: i32) -> Result<(), i32> {
        match unsafe { ffi::ms_comm_init_nccl(self.ctx, id.as_ptr(), rank, world) } { ffi::MS_OK => Ok(()), rc => Err(rc) }
    }
    pub fn unique_id() -> Result<[u8; 128], i32> {
        let mut id = [0u8; 128];
        match unsafe { ffi::ms_comm_unique_id(id.as_mut_ptr()) } { ffi::MS_OK => Ok(id), rc => Err(rc) }
    }
}
impl Drop for Gpu {
    fn drop(&mut self) {
        unsafe { ffi::ms_ctx_destroy(self.ctx) }
    }
}

/// Canonical integer of a one-limb prime-field element (Goldilocks: u64, BabyBear: fits u32).
pub(crate) fn canonical<F: PrimeField>(x: &F) -> u64 {
    x.into_bigint().as_ref()[0]
}

/// The T x W scalar matrix of the AIR's transition closures (src/air.rs:61,119): only constraints LINEAR in the trace
/// polynomials are provable by the reference (the `assert_eq!(rest, zero)` at src/starks.rs:119 demands deg < N), so every
/// closure is a row of scalars; it is recovered by probing the closure with the unit polynomials e_w, and checked on a
/// generic probe.  Same procedure as ministark_b200/air.py `TraceTable.linear_matrix`.
pub(crate) fn linear_matrix<F: PrimeField + ark_ff::FftField>(trace: &TraceTable<F>) -> Result<Vec<u64>, ProverError> {
    let w = trace.width();
    let zero = DensePolynomial::<F>::from_coefficients_vec(vec![]);
    let one = DensePolynomial::<F>::from_coefficients_vec(vec![F::ONE]);
    let mut m = Vec::with_capacity(trace.transition_constrains().len() * w);
    for f in trace.transition_constrains() {
        assert!(f(&vec![zero.clone(); w]).coeffs.is_empty(), "transition constraint with an additive term: not expressible as a T x W matrix");
        for j in 0..w {
            let mut probe = vec![zero.clone(); w];
            probe[j] = one.clone();
            let r = f(&probe);
            assert!(r.coeffs.len() <= 1, "transition constraint multiplies by a non-constant polynomial (src/starks.rs:119 would panic)");
            m.push(r.coeffs.first().map(canonical).unwrap_or(0));
        }
    }
    Ok(m)
}
