//! Seams 2 and 3: `Rep3CoGroth16::prove` / `ShamirCoGroth16::prove` /
//! `prove_with_shamir_bridge` / `Groth16::plain_prove` with the same signatures as co-groth16
//! (co-circom/co-groth16/src/groth16.rs:360-379, 394-417, 439-463, 484-490), executed by the library:
//! witness map + five MSMs on this party's GPU, then the reference's two network legs over the caller's
//! `mpc_net::Network` through the callback transport.
use ark_bn254::{Bn254, Fr};
use ark_groth16::{Proof, ProvingKey};
use co_circom_types::{Rep3SharedWitness, ShamirSharedWitness, SharedWitness};
use co_groth16::ConstraintMatrices;
use cosnarks_gpu_sys as sys;
use mpc_net::Network;
use std::os::raw::{c_int, c_void};

mod key;
pub use key::GpuProvingKey;

/// `&N: Network` as a `cs_net` (id / send / recv).  mpc-net's `send` queues and `recv` blocks -- the contract of
/// `cs_net_callbacks`.
struct NetAdapter<'a, N: Network> { net: &'a N, h: *mut sys::cs_net }

unsafe extern "C" fn send_cb<N: Network>(u: *mut c_void, to: c_int, data: *const c_void, bytes: usize) -> c_int {
    let net = &*(u as *const N);
    let buf = std::slice::from_raw_parts(data as *const u8, bytes);
    if net.send(to as usize, buf).is_ok() { 0 } else { -1 }
}
unsafe extern "C" fn recv_cb<N: Network>(u: *mut c_void, from: c_int, data: *mut c_void, bytes: usize) -> c_int {
    let net = &*(u as *const N);
    match net.recv(from as usize) {
        Ok(v) if v.len() == bytes => { std::ptr::copy_nonoverlapping(v.as_ptr(), data as *mut u8, bytes); 0 }
        Ok(_) => -2,
        Err(_) => -1,
    }
}

impl<'a, N: Network> NetAdapter<'a, N> {
    fn new(net: &'a N, n_parties: usize) -> eyre::Result<Self> {
        let cb = sys::cs_net_callbacks { user: net as *const N as *mut c_void, send: send_cb::<N>, recv: recv_cb::<N> };
        let mut h = std::ptr::null_mut();
        sys::check(unsafe { sys::cs_net_from_callbacks(net.id() as c_int, n_parties as c_int, &cb, &mut h) })
            .map_err(|e| eyre::eyre!(e))?;
        Ok(Self { net, h })
    }
}
impl<N: Network> Drop for NetAdapter<'_, N> { fn drop(&mut self) { unsafe { sys::cs_net_free(self.h) } } }

fn proof_from(a: &[u64; 8], b: &[u64; 16], c: &[u64; 8]) -> Proof<Bn254> { key::proof_from_limbs(a, b, c) }

pub struct Groth16;
pub struct Rep3CoGroth16;
pub struct ShamirCoGroth16;

impl Groth16 {
    /// `Groth16::plain_prove::<R>(pkey, matrices, witness)` (groth16.rs:484-490)
    pub fn plain_prove(pk: &GpuProvingKey, witness: SharedWitness<Fr, Fr>) -> eyre::Result<Proof<Bn254>> {
        pk.check_lengths(witness.public_inputs.len(), witness.witness.len())?;
        let (r, s) = (key::fr_rand()?, key::fr_rand()?); // PlainGroth16Driver::rand (mpc/plain.rs:23-26)
        let (mut a, mut b, mut c) = ([0u64; 8], [0u64; 16], [0u64; 8]);
        sys::check(unsafe {
            sys::cs_groth16_prove_plain(pk.ctx(), pk.ptr(), witness.public_inputs.as_ptr().cast(), witness.witness.as_ptr().cast(),
                                        r.as_ptr(), s.as_ptr(), a.as_mut_ptr(), b.as_mut_ptr(), c.as_mut_ptr())
        }).map_err(|e| eyre::eyre!(e))?;
        Ok(proof_from(&a, &b, &c))
    }
}

impl Rep3CoGroth16 {
    /// `Rep3CoGroth16::prove::<N, CircomReduction>(net0, net1, &pkey, &matrices, witness)` (groth16.rs:360-379).
    /// `Rep3State::new(net0)` + `fork` happen inside (`cs_rep3_state_create`: OS-entropy seed, `reshare`).
    pub fn prove<N: Network>(net0: &N, net1: &N, pk: &GpuProvingKey, _matrices: &ConstraintMatrices<Fr>,
                             witness: Rep3SharedWitness<Fr>) -> eyre::Result<Proof<Bn254>> {
        pk.check_lengths(witness.public_inputs.len(), witness.witness.len())?;
        let (n0, n1) = (NetAdapter::new(net0, 3)?, NetAdapter::new(net1, 3)?);
        let mut state = std::ptr::null_mut();
        sys::check(unsafe { sys::cs_rep3_state_create(n0.h, &mut state) }).map_err(|e| eyre::eyre!(e))?;
        let (mut a, mut b, mut c) = ([0u64; 8], [0u64; 16], [0u64; 8]);
        // Rep3PrimeFieldShare<Fr> { a, b } is two consecutive [u64; 4]: the share vector crosses as it lies
        let rc = unsafe {
            sys::cs_groth16_rep3_prove(pk.ctx(), pk.ptr(), n0.h, n1.h, state, witness.public_inputs.as_ptr().cast(),
                                       witness.witness.as_ptr().cast(), std::ptr::null(), a.as_mut_ptr(), b.as_mut_ptr(),
                                       c.as_mut_ptr(), std::ptr::null_mut())
        };
        unsafe { sys::cs_rep3_state_free(state) };
        sys::check(rc).map_err(|e| eyre::eyre!(e))?;
        Ok(proof_from(&a, &b, &c))
    }

    /// `prove_with_shamir_bridge` (groth16.rs:394-417): local translation to Shamir(t = 1) + Shamir prover
    pub fn prove_with_shamir_bridge<N: Network>(net0: &N, net1: &N, pk: &GpuProvingKey, _matrices: &ConstraintMatrices<Fr>,
                                                witness: Rep3SharedWitness<Fr>) -> eyre::Result<Proof<Bn254>> {
        pk.check_lengths(witness.public_inputs.len(), witness.witness.len())?;
        let (n0, n1) = (NetAdapter::new(net0, 3)?, NetAdapter::new(net1, 3)?);
        let (mut a, mut b, mut c) = ([0u64; 8], [0u64; 16], [0u64; 8]);
        sys::check(unsafe {
            sys::cs_groth16_prove_with_shamir_bridge(pk.ctx(), pk.ptr(), n0.h, n1.h, witness.public_inputs.as_ptr().cast(),
                                                     witness.witness.as_ptr().cast(), a.as_mut_ptr(), b.as_mut_ptr(),
                                                     c.as_mut_ptr(), std::ptr::null_mut())
        }).map_err(|e| eyre::eyre!(e))?;
        Ok(proof_from(&a, &b, &c))
    }
}

impl ShamirCoGroth16 {
    /// `ShamirCoGroth16::prove` (groth16.rs:439-463)
    pub fn prove<N: Network>(net0: &N, net1: &N, num_parties: usize, threshold: usize, pk: &GpuProvingKey,
                             _matrices: &ConstraintMatrices<Fr>, witness: ShamirSharedWitness<Fr>) -> eyre::Result<Proof<Bn254>> {
        pk.check_lengths(witness.public_inputs.len(), witness.witness.len())?;
        let (n0, n1) = (NetAdapter::new(net0, num_parties)?, NetAdapter::new(net1, num_parties)?);
        let (mut a, mut b, mut c) = ([0u64; 8], [0u64; 16], [0u64; 8]);
        sys::check(unsafe {
            sys::cs_groth16_shamir_prove(pk.ctx(), pk.ptr(), n0.h, n1.h, num_parties as c_int, threshold as c_int,
                                         witness.public_inputs.as_ptr().cast(), witness.witness.as_ptr().cast(),
                                         a.as_mut_ptr(), b.as_mut_ptr(), c.as_mut_ptr(), std::ptr::null_mut())
        }).map_err(|e| eyre::eyre!(e))?;
        Ok(proof_from(&a, &b, &c))
    }
}
