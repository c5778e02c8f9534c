//! `Stark::prove` / `Stark::verify` on libministark.so -- the bodies that replace src/starks.rs:59-169 and :171-235 under
//! feature `b200`.  Placed as a child module of `starks` in the reference crate (`#[cfg(feature = "b200")] mod starks_gpu;`
//! inside src/starks.rs), so that it can build the crate's `StarkProof` (its fields are private to `starks`).
//! NOT COMPILED in this repository's image (no cargo / rustc).
use super::{canonical, ffi, linear_matrix, Gpu, GpuField};
use crate::air::{Constrains, Provable};
use crate::error::{ProverError, VerifierError};
use crate::starks::{Stark, StarkProof};
use digest::core_api::BlockSizeUser;
use digest::{Digest, FixedOutputReset};

fn params<D, F>(s: &Stark<D, F>) -> ffi::ms_stark_params
where
    F: GpuField,
    D: Digest + FixedOutputReset + BlockSizeUser + Clone,
{
    let c = s.config(); // &StarkConfig: security_bits, blowup_factor, steps, merkle_config.leafs_per_node (src/starks.rs:238-257)
    ffi::ms_stark_params {
        security_bits: c.security_bits as u64,
        blowup_factor: c.blowup_factor as u64,
        steps: c.steps as u64,
        trace_columns: c.merkle_config.leafs_per_node as u64,
        inner_children: 2, // src/starks.rs:299
    }
}

impl<D, F> Stark<D, F>
where
    F: GpuField,
    D: Digest + FixedOutputReset + BlockSizeUser + Clone,
{
    /// Stark::prove (src/starks.rs:59-169) with everything behind `air.trace(&witness)` on the GPU.
    pub fn prove_gpu<T, AIR: Provable<T, F::Base>>(&self, gpu: &Gpu, air: AIR, witness: T) -> Result<StarkProof<D, F>, ProverError> {
        let trace = air.trace(&witness); // USER CODE, unchanged (src/starks.rs:68)
        let (n, w) = (trace.len(), trace.width());
        let elem = if F::FIELD_ID == ffi::MS_FIELD_GOLDILOCKS { 8 } else { 4 };
        // canonical row-major trace in the element width of the field (u64 / u32)
        let mut rows = vec![0u8; n * w * elem];
        for (i, x) in trace.data().iter().enumerate() {
            rows[i * elem..(i + 1) * elem].copy_from_slice(&canonical(x).to_le_bytes()[..elem]);
        }
        let matrix64 = linear_matrix(&trace)?;
        let t = matrix64.len() / w;
        let mut matrix = vec![0u8; matrix64.len() * elem];
        for (i, v) in matrix64.iter().enumerate() {
            matrix[i * elem..(i + 1) * elem].copy_from_slice(&v.to_le_bytes()[..elem]);
        }
        let p = params(self);
        let mut len = unsafe { ffi::ms_stark_proof_bound(F::FIELD_ID, &p, n as u64, (w + t) as u64) };
        let mut buf = vec![0u8; len as usize];
        let rc = unsafe {
            ffi::ms_stark_prove(gpu.ctx, &p, rows.as_ptr().cast(), n as u64, w as u64, matrix.as_ptr().cast(), t as u64, buf.as_mut_ptr(), &mut len)
        };
        match rc {
            ffi::MS_OK => Ok(StarkProof::from_dump(&buf[..len as usize])), // field order of src/starks.rs:21-28, DESIGN.md "Proof bytes"
            // shape violations the reference turns into panics (src/merkle.rs:95,99-104; src/air.rs:23; src/starks.rs:84-85,119)
            ffi::MS_ERR_BAD_SHAPE | ffi::MS_ERR_QUOTIENT_NONZERO => panic!("{}", gpu.last_error()),
            ffi::MS_ERR_LEAF_NOT_FOUND => Err(ProverError::LeafNotFound { msg: "leaf is not included in the tree" }), // src/error.rs:13-16
            _ => Err(ProverError::from_gpu(rc, gpu.last_error())), // transcript / CUDA / NCCL errors (src/error.rs:5-8)
        }
    }

    /// Stark::verify (src/starks.rs:171-235): the Constrains polynomials are uploaded and re-evaluated on the device.
    pub fn verify_gpu(&self, gpu: &Gpu, constrains: Constrains<F::Base>, proof: &StarkProof<D, F>) -> Result<bool, VerifierError> {
        let polys = constrains.get_polynomials();
        let n = self.config().steps.next_power_of_two().max(polys.iter().map(|p| p.coeffs.len()).max().unwrap_or(1).next_power_of_two());
        let elem = if F::FIELD_ID == ffi::MS_FIELD_GOLDILOCKS { 8 } else { 4 };
        let mut host = vec![0u8; polys.len() * n * elem];
        for (c, p) in polys.iter().enumerate() {
            for (m, x) in p.coeffs.iter().enumerate() {
                let at = (c * n + m) * elem;
                host[at..at + elem].copy_from_slice(&canonical(x).to_le_bytes()[..elem]);
            }
        }
        let dump = proof.to_dump();
        let (mut d_ptr, mut ok, mut line) = (std::ptr::null_mut(), 0i32, 0i32);
        unsafe {
            assert_eq!(ffi::ms_dev_alloc(gpu.ctx, host.len(), &mut d_ptr), ffi::MS_OK);
            assert_eq!(ffi::ms_h2d(gpu.ctx, d_ptr, host.as_ptr().cast(), host.len()), ffi::MS_OK);
            let rc = ffi::ms_stark_verify(gpu.ctx, &params(self), d_ptr, n as u64, n as u64, polys.len() as u64, dump.as_ptr(), dump.len() as u64, 0, &mut ok, &mut line);
            ffi::ms_dev_free(gpu.ctx, d_ptr);
            assert_eq!(rc, ffi::MS_OK, "{}", gpu.last_error());
        }
        // the reference's checks are `assert!`s: a failed check panics there, here it names the line
        assert!(ok == 1, "Stark::verify: check at src line {line} failed");
        Ok(true)
    }
}
