//! `Groth16B200`: ark-snark's `SNARK` trait over the B200 backend (include/b200snark.h).
//! SOURCE ONLY -- never compiled here (no Rust toolchain in the build image).  See INTEGRATION.md.
use std::os::raw::c_void;

use ark_ec::pairing::Pairing;
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

pub struct Groth16B200<E: Pairing>(core::marker::PhantomData<E>);

/// Device handles for one (proving key, circuit shape): matrices and key are witness independent.
pub struct Resident { pub ctx: *mut B2sCtx, pub pk: *mut B2sPk, pub mat: *mut B2sR1cs }

impl<E: Pairing> Groth16B200<E> {
    /// Upload the matrices and the key (packing `Affine { x, y, infinity }` into x||y, zeros for infinity).
    pub fn make_resident(_pk: &ProvingKey<E>, _mats: &[Matrix<E::ScalarField>], _n_inst: usize, _n_wit: usize)
        -> Result<Resident, B200Error> {
        unimplemented!("pack points field-wise, call b2s_r1cs_upload / b2s_pk_upload; see INTEGRATION.md section 2")
    }
}

impl<E: Pairing> SNARK<E::ScalarField> for Groth16B200<E> {
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
        let h = Self::make_resident(pk, &mats, zi.len(), zw.len())?;   // cached per circuit in a real adapter
        let g1 = 2 * core::mem::size_of::<E::BaseField>();
        let (mut a, mut b, mut c) = (vec![0u8; g1], vec![0u8; 2 * g1], vec![0u8; g1]);
        let st = unsafe {
            b2s_groth16_prove(h.ctx, h.pk, h.mat, zi.as_ptr().cast(), zw.as_ptr().cast(), (&r as *const E::ScalarField).cast(),
                              (&s as *const E::ScalarField).cast(), a.as_mut_ptr().cast(), b.as_mut_ptr().cast(), c.as_mut_ptr().cast())
        };
        if st != 0 { return Err(B200Error::from_status(st)); }
        unimplemented!("unpack a, b, c (x||y Montgomery limbs, zeros = infinity) into Proof {{ a, b, c }}")
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

impl<E: Pairing> CircuitSpecificSetupSNARK<E::ScalarField> for Groth16B200<E> {}
