//! `ark_groth16::ProvingKey<Bn254>` + `ConstraintMatrices<Fr>` -> device-resident `cs_groth16_pk` (uploaded once,
//! kept next to the key), plus the small conversions at the boundary.
use ark_bn254::{Bn254, Fq, Fq2, Fr, G1Affine, G2Affine};
use ark_ec::AffineRepr;
use ark_groth16::{Proof, ProvingKey};
use co_groth16::ConstraintMatrices;
use cosnarks_gpu_sys as sys;

pub struct GpuProvingKey { ctx: *mut sys::cs_ctx, pk: *mut sys::cs_groth16_pk, ni: usize, nw: usize }
unsafe impl Send for GpuProvingKey {}
unsafe impl Sync for GpuProvingKey {}

fn g1(p: &G1Affine) -> [u64; 8] {
    let mut o = [0u64; 8];
    if let Some((x, y)) = p.xy() { o[..4].copy_from_slice(&x.0 .0); o[4..].copy_from_slice(&y.0 .0); }
    o
}
fn g2(p: &G2Affine) -> [u64; 16] {
    let mut o = [0u64; 16];
    if let Some((x, y)) = p.xy() {
        o[..4].copy_from_slice(&x.c0.0 .0); o[4..8].copy_from_slice(&x.c1.0 .0);
        o[8..12].copy_from_slice(&y.c0.0 .0); o[12..].copy_from_slice(&y.c1.0 .0);
    }
    o
}
fn csr(rows: &[Vec<(Fr, usize)>]) -> (Vec<u32>, Vec<u32>, Vec<u64>) {
    let (mut rp, mut col, mut cf) = (vec![0u32], vec![], vec![]);
    for row in rows {
        for (c, i) in row { col.push(*i as u32); cf.extend_from_slice(&c.0 .0); }
        rp.push(col.len() as u32);
    }
    (rp, col, cf)
}

impl GpuProvingKey {
    pub fn new(device: i32, pkey: &ProvingKey<Bn254>, m: &ConstraintMatrices<Fr>) -> eyre::Result<Self> {
        let mut ctx = std::ptr::null_mut();
        sys::check(unsafe { sys::cs_ctx_create(device, std::ptr::null_mut(), &mut ctx) }).map_err(|e| eyre::eyre!(e))?;
        let (a, b) = (csr(&m.a), csr(&m.b));
        let flat1 = |v: &[G1Affine]| v.iter().flat_map(g1).collect::<Vec<u64>>();
        let flat2 = |v: &[G2Affine]| v.iter().flat_map(g2).collect::<Vec<u64>>();
        let (aq, b1q, b2q, lq, hq) = (flat1(&pkey.a_query), flat1(&pkey.b_g1_query), flat2(&pkey.b_g2_query), flat1(&pkey.l_query), flat1(&pkey.h_query));
        let (al, be1, be2, de1, de2) = (g1(&pkey.vk.alpha_g1), g1(&pkey.beta_g1), g2(&pkey.vk.beta_g2), g1(&pkey.delta_g1), g2(&pkey.vk.delta_g2));
        let d = sys::cs_groth16_key_desc {
            curve: sys::CS_BN254, num_constraints: m.num_constraints, num_instance_variables: m.num_instance_variables,
            num_witness_variables: m.num_witness_variables,
            a_row_ptr: a.0.as_ptr(), a_col: a.1.as_ptr(), a_coeff: a.2.as_ptr(), a_nnz: a.1.len(),
            b_row_ptr: b.0.as_ptr(), b_col: b.1.as_ptr(), b_coeff: b.2.as_ptr(), b_nnz: b.1.len(),
            c_row_ptr: std::ptr::null(), c_col: std::ptr::null(), c_coeff: std::ptr::null(), c_nnz: 0,
            alpha_g1: al.as_ptr(), beta_g1: be1.as_ptr(), beta_g2: be2.as_ptr(), delta_g1: de1.as_ptr(), delta_g2: de2.as_ptr(),
            a_query: aq.as_ptr(), a_query_len: pkey.a_query.len(), b_g1_query: b1q.as_ptr(), b_g1_query_len: pkey.b_g1_query.len(),
            b_g2_query: b2q.as_ptr(), b_g2_query_len: pkey.b_g2_query.len(), l_query: lq.as_ptr(), l_query_len: pkey.l_query.len(),
            h_query: hq.as_ptr(), h_query_len: pkey.h_query.len(), window_bits: 0,
        };
        let mut pk = std::ptr::null_mut();
        sys::check(unsafe { sys::cs_groth16_pk_create(ctx, &d, &mut pk) }).map_err(|e| eyre::eyre!(e))?;
        Ok(Self { ctx, pk, ni: m.num_instance_variables, nw: m.num_witness_variables })
    }
    pub(crate) fn ctx(&self) -> *mut sys::cs_ctx { self.ctx }
    pub(crate) fn ptr(&self) -> *mut sys::cs_groth16_pk { self.pk }
    /// the length checks of prove_inner (groth16.rs:134-149), same messages
    pub(crate) fn check_lengths(&self, n_pub: usize, n_wit: usize) -> eyre::Result<()> {
        if n_pub != self.ni {
            eyre::bail!("amount of public inputs does not match with provided constraint system! Expected {}, but got {}", self.ni, n_pub)
        }
        if n_wit != self.nw {
            eyre::bail!("amount of private witness variables does not match with provided constraint system! Expected {}, but got {}", self.nw, n_wit)
        }
        Ok(())
    }
}
impl Drop for GpuProvingKey {
    fn drop(&mut self) { unsafe { sys::cs_groth16_pk_free(self.pk); sys::cs_ctx_destroy(self.ctx) } }
}

pub(crate) fn proof_from_limbs(a: &[u64; 8], b: &[u64; 16], c: &[u64; 8]) -> Proof<Bn254> {
    let fq = |l: &[u64]| Fq::new_unchecked(ark_ff::BigInt([l[0], l[1], l[2], l[3]]));
    let p1 = |l: &[u64; 8]| if l.iter().all(|x| *x == 0) { G1Affine::zero() } else { G1Affine::new_unchecked(fq(&l[..4]), fq(&l[4..])) };
    let p2 = |l: &[u64; 16]| if l.iter().all(|x| *x == 0) { G2Affine::zero() } else {
        G2Affine::new_unchecked(Fq2::new(fq(&l[..4]), fq(&l[4..8])), Fq2::new(fq(&l[8..12]), fq(&l[12..])))
    };
    Proof { a: p1(a), b: p2(b), c: p1(c) }
}

/// uniform Fr from the OS entropy pool (rejection sampling), as Montgomery limbs
pub(crate) fn fr_rand() -> eyre::Result<[u64; 4]> {
    const R: [u64; 4] = [0x43e1f593f0000001, 0x2833e84879b97091, 0xb85045b68181585d, 0x30644e72e131a029];
    loop {
        let mut v = [0u64; 4];
        sys::check(unsafe { sys::cs_os_random(v.as_mut_ptr().cast(), 32) }).map_err(|e| eyre::eyre!(e))?;
        v[3] &= (1u64 << 62) - 1;
        for i in (0..4).rev() {
            if v[i] < R[i] { return Ok(v); }
            if v[i] > R[i] { break; }
        }
    }
}
