//! `Groth16B200`: ark-snark's `SNARK` trait over the B200 backend (include/b200snark.h).
//! SOURCE ONLY -- never compiled here (no Rust toolchain in the build image).  See INTEGRATION.md.
use std::os::raw::c_void;

use ark_ec::{pairing::Pairing, AffineRepr};
use ark_groth16::{Groth16, PreparedVerifyingKey, Proof, ProvingKey, VerifyingKey};
use ark_relations::gr1cs::{
    ConstraintSynthesizer, ConstraintSystem, Matrix, OptimizationGoal, SynthesisError, R1CS_PREDICATE_LABEL,
};
use ark_snark::{CircuitSpecificSetupSNARK, SNARK};
use ark_std::{rand::{CryptoRng, RngCore}, UniformRand};

#[repr(C)] pub struct B2sCtx { _p: [u8; 0] }
#[repr(C)] pub struct B2sR1cs { _p: [u8; 0] }
#[repr(C)] pub struct B2sPk { _p: [u8; 0] }

#[repr(C)]
pub struct B2sPkDesc {
    pub n_instance: u64, pub n_witness: u64, pub domain_size: u64,
    pub alpha_g1: *const c_void, pub beta_g1: *const c_void, pub delta_g1: *const c_void,
    pub beta_g2: *const c_void, pub delta_g2: *const c_void,
    pub a_query: *const c_void, pub a_off: u64, pub a_len: u64,
    pub b_g1_query: *const c_void, pub b1_off: u64, pub b1_len: u64,
    pub b_g2_query: *const c_void, pub b2_off: u64, pub b2_len: u64,
    pub h_query: *const c_void, pub h_off: u64, pub h_len: u64,
    pub l_query: *const c_void, pub l_off: u64, pub l_len: u64,
}

extern "C" {
    pub fn b2s_ctx_create(curve_id: i32, device: i32, out: *mut *mut B2sCtx) -> i32;
    pub fn b2s_ctx_destroy(ctx: *mut B2sCtx);
    pub fn b2s_last_error(ctx: *const B2sCtx) -> *const std::os::raw::c_char;
    pub fn b2s_r1cs_upload(ctx: *mut B2sCtx, n_rows: u64, n_inst: u64, n_wit: u64, row_ptr: *const *const u64,
                           col: *const *const u32, coeff: *const *const c_void, out: *mut *mut B2sR1cs) -> i32;
    pub fn b2s_r1cs_free(ctx: *mut B2sCtx, m: *mut B2sR1cs);
    pub fn b2s_pk_upload(ctx: *mut B2sCtx, desc: *const B2sPkDesc, mem: i32, out: *mut *mut B2sPk) -> i32;
    pub fn b2s_pk_free(ctx: *mut B2sCtx, pk: *mut B2sPk);
    pub fn b2s_groth16_prove(ctx: *mut B2sCtx, pk: *const B2sPk, m: *const B2sR1cs, z_inst: *const c_void,
                             z_wit: *const c_void, r: *const c_void, s: *const c_void, out_a: *mut c_void,
                             out_b: *mut c_void, out_c: *mut c_void) -> i32;
    // multi-GPU group (one rank per GPU, NCCL communicator inside the library)
    pub fn b2s_group_unique_id(out: *mut u8) -> i32; // 128 bytes
    pub fn b2s_group_create(ctx: *mut B2sCtx, id: *const u8, rank: i32, world: i32, out: *mut *mut B2sGroup) -> i32;
    pub fn b2s_group_destroy(group: *mut B2sGroup);
    pub fn b2s_groth16_prove_group(group: *mut B2sGroup, pk_shard: *const B2sPk, m: *const B2sR1cs, z_inst: *const c_void,
                                   z_wit: *const c_void, r: *const c_void, s: *const c_void, out_a: *mut c_void,
                                   out_b: *mut c_void, out_c: *mut c_void) -> i32;
    // CanonicalSerialize of the key types (snark/src/lib.rs:25-31)
    pub fn b2s_vk_serialized_size(ctx: *const B2sCtx, n_gamma_abc: u64, compressed: i32) -> u64;
    pub fn b2s_vk_serialize(ctx: *mut B2sCtx, alpha_g1: *const c_void, beta_g2: *const c_void, gamma_g2: *const c_void,
                            delta_g2: *const c_void, gamma_abc_g1: *const c_void, n_gamma_abc: u64, compressed: i32,
                            out: *mut u8, cap: u64) -> i32;
    pub fn b2s_pk_serialized_size(ctx: *const B2sCtx, pk: *const B2sPk, vk_len: u64, compressed: i32) -> u64;
    pub fn b2s_pk_serialize(ctx: *mut B2sCtx, pk: *const B2sPk, vk_bytes: *const u8, vk_len: u64, compressed: i32,
                            out: *mut u8, cap: u64) -> i32;
    // universal-setup schemes (UniversalSetupSNARK, snark/src/lib.rs:107-133): the seams a polynomial-commitment /
    // evaluation-domain backend binds.  mem: 0 host, 1 device; s, c, z: one Montgomery Fr on the host
    pub fn b2s_ntt(ctx: *mut B2sCtx, data: *mut c_void, log_n: u32, inverse: i32, coset: i32, mem: i32) -> i32;
    pub fn b2s_msm_g1(ctx: *mut B2sCtx, bases: *const c_void, scalars: *const c_void, n: u64, scalars_mont: i32, mem: i32,
                      out_affine: *mut c_void) -> i32;
    pub fn b2s_fixed_base_g1(ctx: *mut B2sCtx, scalars: *const c_void, n: u64, scalars_mont: i32, mem: i32, out: *mut c_void) -> i32;
    pub fn b2s_poly_op(ctx: *mut B2sCtx, op: i32, a: *const c_void, b: *const c_void, s: *const c_void, out: *mut c_void,
                       n: u64, mem: i32) -> i32;
    pub fn b2s_poly_geom(ctx: *mut B2sCtx, c: *const c_void, s: *const c_void, n: u64, mem: i32, out: *mut c_void) -> i32;
    pub fn b2s_poly_eval(ctx: *mut B2sCtx, coeffs: *const c_void, n: u64, z: *const c_void, mem: i32, out: *mut c_void) -> i32;
}
#[repr(C)]
pub struct B2sGroup {
    _p: [u8; 0],
}

#[derive(Debug)]
pub enum B200Error { Synthesis(SynthesisError), Backend(i32) }
impl From<SynthesisError> for B200Error { fn from(e: SynthesisError) -> Self { B200Error::Synthesis(e) } }
impl core::fmt::Display for B200Error {
    fn fmt(&self, f: &mut core::fmt::Formatter<'_>) -> core::fmt::Result { write!(f, "{self:?}") }
}
impl ark_std::error::Error for B200Error {}
impl B200Error {
    /// status codes 1..7 mirror SynthesisError (relations/src/utils/error.rs:5-21)
    pub fn from_status(st: i32) -> Self {
        match st {
            1 => SynthesisError::MissingCS.into(),
            2 => SynthesisError::AssignmentMissing.into(),
            3 => SynthesisError::DivisionByZero.into(),
            4 => SynthesisError::Unsatisfiable.into(),
            5 => SynthesisError::PolynomialDegreeTooLarge.into(),
            6 => SynthesisError::UnexpectedIdentity.into(),
            7 => SynthesisError::MalformedVerifyingKey.into(),
            s => B200Error::Backend(s),
        }
    }
}

/// `Matrix<F>` (relations/src/utils/matrix.rs:4) -> CSR, once per circuit.
pub fn to_csr<F: Copy>(m: &Matrix<F>) -> (Vec<u64>, Vec<u32>, Vec<F>) {
    let mut rp = Vec::with_capacity(m.len() + 1);
    let (mut col, mut co) = (Vec::new(), Vec::new());
    rp.push(0u64);
    for row in m {
        for (c, j) in row { col.push(*j as u32); co.push(*c); }
        rp.push(col.len() as u64);
    }
    (rp, col, co)
}

/// The in-memory bytes of a field element.  ark-ff's `Fp<MontBackend<_, N>>` is `BigInt<N>` (N little-endian u64
/// limbs, Montgomery form) plus a zero-sized marker, and `Fp2` is `{ c0, c1 }` of those -- exactly the C-ABI layout.
fn raw<T>(t: &T) -> &[u8] { unsafe { core::slice::from_raw_parts((t as *const T).cast::<u8>(), core::mem::size_of::<T>()) } }

/// `Affine {{ x, y, infinity }}` -> x || y, all-zero bytes for the point at infinity (include/b200snark.h, "Layouts").
pub fn pack_points<G: AffineRepr>(pts: &[G]) -> Vec<u8> {
    let w = core::mem::size_of::<G::BaseField>();
    let mut out = vec![0u8; 2 * w * pts.len()];
    for (i, p) in pts.iter().enumerate() {
        if let Some((x, y)) = p.xy() {
            out[2 * w * i..2 * w * i + w].copy_from_slice(raw(&x));
            out[2 * w * i + w..2 * w * (i + 1)].copy_from_slice(raw(&y));
        }
    }
    out
}

/// x || y (Montgomery limbs; zeros = infinity) -> `Affine`.  The coordinates come from the prover, so they are on the
/// curve by construction: `new_unchecked`.
pub fn unpack_point<G: AffineRepr>(bytes: &[u8]) -> G where G::BaseField: Copy {
    let w = core::mem::size_of::<G::BaseField>();
    if bytes.iter().all(|b| *b == 0) { return G::zero(); }
    let read = |b: &[u8]| -> G::BaseField { unsafe { core::ptr::read_unaligned(b.as_ptr().cast::<G::BaseField>()) } };
    G::new_unchecked(read(&bytes[..w]), read(&bytes[w..2 * w]))
}

fn check(ctx: *mut B2sCtx, st: i32) -> Result<(), B200Error> {
    let _ = ctx;   // b2s_last_error(ctx) carries the message for logs
    if st == 0 { Ok(()) } else { Err(B200Error::from_status(st)) }
}

pub struct Groth16B200<E: Pairing>(core::marker::PhantomData<E>);

/// Device handles for one (proving key, circuit shape): matrices and key are witness independent.
pub struct Resident { pub ctx: *mut B2sCtx, pub pk: *mut B2sPk, pub mat: *mut B2sR1cs }

impl<E: Pairing> Groth16B200<E> {
    /// Upload the matrices and the key once per (key, circuit shape).  `curve_id`: 0 = BLS12-381, 1 = BN254.
    pub fn make_resident(curve_id: i32, pk: &ProvingKey<E>, mats: &[Matrix<E::ScalarField>], n_inst: usize, n_wit: usize)
        -> Result<Resident, B200Error> {
        let mut ctx: *mut B2sCtx = core::ptr::null_mut();
        check(ctx, unsafe { b2s_ctx_create(curve_id, 0, &mut ctx) })?;
        let csr: Vec<_> = mats.iter().map(to_csr).collect();
        let rp: Vec<*const u64> = csr.iter().map(|m| m.0.as_ptr()).collect();
        let col: Vec<*const u32> = csr.iter().map(|m| m.1.as_ptr()).collect();
        let co: Vec<*const c_void> = csr.iter().map(|m| m.2.as_ptr().cast()).collect();
        let mut mat: *mut B2sR1cs = core::ptr::null_mut();
        check(ctx, unsafe { b2s_r1cs_upload(ctx, mats[0].len() as u64, n_inst as u64, n_wit as u64, rp.as_ptr(), col.as_ptr(), co.as_ptr(), &mut mat) })?;
        let n = (mats[0].len() + n_inst).next_power_of_two() as u64;          // the QAP domain (LibsnarkReduction)
        let (alpha, beta1, delta1) = (pack_points(&[pk.vk.alpha_g1]), pack_points(&[pk.beta_g1]), pack_points(&[pk.delta_g1]));
        let (beta2, delta2) = (pack_points(&[pk.vk.beta_g2]), pack_points(&[pk.vk.delta_g2]));
        let (a, b1, b2, h, l) = (pack_points(&pk.a_query), pack_points(&pk.b_g1_query), pack_points(&pk.b_g2_query),
                                 pack_points(&pk.h_query), pack_points(&pk.l_query));
        let d = B2sPkDesc {
            n_instance: n_inst as u64, n_witness: n_wit as u64, domain_size: n,
            alpha_g1: alpha.as_ptr().cast(), beta_g1: beta1.as_ptr().cast(), delta_g1: delta1.as_ptr().cast(),
            beta_g2: beta2.as_ptr().cast(), delta_g2: delta2.as_ptr().cast(),
            a_query: a.as_ptr().cast(), a_off: 0, a_len: pk.a_query.len() as u64,
            b_g1_query: b1.as_ptr().cast(), b1_off: 0, b1_len: pk.b_g1_query.len() as u64,
            b_g2_query: b2.as_ptr().cast(), b2_off: 0, b2_len: pk.b_g2_query.len() as u64,
            h_query: h.as_ptr().cast(), h_off: 0, h_len: pk.h_query.len() as u64,
            l_query: l.as_ptr().cast(), l_off: 0, l_len: pk.l_query.len() as u64,
        };
        let mut pkh: *mut B2sPk = core::ptr::null_mut();
        check(ctx, unsafe { b2s_pk_upload(ctx, &d, 0 /* B2S_MEM_HOST */, &mut pkh) })?;
        Ok(Resident { ctx, pk: pkh, mat })
    }
}

impl Drop for Resident {
    fn drop(&mut self) { unsafe { b2s_pk_free(self.ctx, self.pk); b2s_r1cs_free(self.ctx, self.mat); b2s_ctx_destroy(self.ctx); } }
}

/// Which curve id the backend should use for `E` (the backend supports the two curves of the north-star).
pub trait B200Curve { const CURVE_ID: i32; }

impl<E: Pairing + B200Curve> SNARK<E::ScalarField> for Groth16B200<E> {
    type ProvingKey = ProvingKey<E>;
    type VerifyingKey = VerifyingKey<E>;
    type Proof = Proof<E>;
    type ProcessedVerifyingKey = PreparedVerifyingKey<E>;
    type Error = B200Error;

    fn circuit_specific_setup<C: ConstraintSynthesizer<E::ScalarField>, R: RngCore + CryptoRng>(
        circuit: C, rng: &mut R,
    ) -> Result<(Self::ProvingKey, Self::VerifyingKey), Self::Error> {
        Groth16::<E>::circuit_specific_setup(circuit, rng).map_err(B200Error::from)
    }

    fn prove<C: ConstraintSynthesizer<E::ScalarField>, R: RngCore + CryptoRng>(
        pk: &Self::ProvingKey, circuit: C, rng: &mut R,
    ) -> Result<Self::Proof, Self::Error> {
        let r = E::ScalarField::rand(rng);
        let s = E::ScalarField::rand(rng);
        let cs = ConstraintSystem::new_ref();
        cs.set_optimization_goal(OptimizationGoal::Constraints);
        circuit.generate_constraints(cs.clone())?;
        cs.finalize();
        let mats = cs.to_matrices()?.remove(R1CS_PREDICATE_LABEL).ok_or(SynthesisError::MissingCS)?;
        let (zi, zw) = (cs.instance_assignment()?, cs.witness_assignment()?);
        let h = Self::make_resident(E::CURVE_ID, pk, &mats, zi.len(), zw.len())?;   // cache per (key, circuit) in a long-lived prover
        let g1 = 2 * core::mem::size_of::<<E::G1Affine as AffineRepr>::BaseField>();
        let (mut a, mut b, mut c) = (vec![0u8; g1], vec![0u8; 2 * g1], vec![0u8; g1]);
        let st = unsafe {
            b2s_groth16_prove(h.ctx, h.pk, h.mat, zi.as_ptr().cast(), zw.as_ptr().cast(), (&r as *const E::ScalarField).cast(),
                              (&s as *const E::ScalarField).cast(), a.as_mut_ptr().cast(), b.as_mut_ptr().cast(), c.as_mut_ptr().cast())
        };
        if st != 0 { return Err(B200Error::from_status(st)); }
        Ok(Proof { a: unpack_point::<E::G1Affine>(&a), b: unpack_point::<E::G2Affine>(&b), c: unpack_point::<E::G1Affine>(&c) })
    }

    fn process_vk(vk: &Self::VerifyingKey) -> Result<Self::ProcessedVerifyingKey, Self::Error> {
        Ok(ark_groth16::prepare_verifying_key(vk))
    }

    fn verify_with_processed_vk(
        pvk: &Self::ProcessedVerifyingKey, x: &[E::ScalarField], proof: &Self::Proof,
    ) -> Result<bool, Self::Error> {
        Groth16::<E>::verify_with_processed_vk(pvk, x, proof).map_err(B200Error::from)
    }
}

impl<E: Pairing + B200Curve> CircuitSpecificSetupSNARK<E::ScalarField> for Groth16B200<E> {}

// e.g. in the application:  impl B200Curve for ark_bls12_381::Bls12_381 { const CURVE_ID: i32 = 0; }
//                           impl B200Curve for ark_bn254::Bn254 { const CURVE_ID: i32 = 1; }
