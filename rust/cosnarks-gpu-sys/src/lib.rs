//! Raw declarations of the C ABI in `include/cosnarks_gpu.h`.  Field elements are `[u64; N]` Montgomery limbs
//! (byte-identical to arkworks' `Fp<MontBackend<_, N>>`), points are packed `x || y` limb arrays with the
//! all-zero encoding for infinity, Rep3 shares are `a || b`.  Every function returns 0 or a negative code;
//! `cs_last_error()` holds the message for the calling thread.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_uint, c_void};

macro_rules! opaque { ($($n:ident),*) => { $( #[repr(C)] pub struct $n { _p: [u8; 0] } )* } }
opaque!(cs_ctx, cs_bases, cs_domain, cs_groth16_pk, cs_plonk_pk, cs_net, cs_rep3_state, cs_shamir_state);

pub const CS_BN254: c_int = 0;
pub const CS_BLS12_381: c_int = 1;
pub const CS_G1: c_int = 0;
pub const CS_G2: c_int = 1;
pub const CS_PLAIN: c_int = 0;
pub const CS_REP3: c_int = 1;

#[repr(C)]
pub struct cs_net_callbacks {
    pub user: *mut c_void,
    pub send: unsafe extern "C" fn(user: *mut c_void, to: c_int, data: *const c_void, bytes: usize) -> c_int,
    pub recv: unsafe extern "C" fn(user: *mut c_void, from: c_int, data: *mut c_void, bytes: usize) -> c_int,
}

#[repr(C)]
pub struct cs_rep3_prf {
    pub seed1: [u8; 32],
    pub word_pos1: u64,
    pub seed2: [u8; 32],
    pub word_pos2: u64,
    pub rounds: c_uint,
}

#[repr(C)]
pub struct cs_groth16_key_desc {
    pub curve: c_int,
    pub num_constraints: usize,
    pub num_instance_variables: usize,
    pub num_witness_variables: usize,
    pub a_row_ptr: *const u32, pub a_col: *const u32, pub a_coeff: *const u64, pub a_nnz: usize,
    pub b_row_ptr: *const u32, pub b_col: *const u32, pub b_coeff: *const u64, pub b_nnz: usize,
    pub c_row_ptr: *const u32, pub c_col: *const u32, pub c_coeff: *const u64, pub c_nnz: usize,
    pub alpha_g1: *const u64, pub beta_g1: *const u64, pub beta_g2: *const u64,
    pub delta_g1: *const u64, pub delta_g2: *const u64,
    pub a_query: *const u64, pub a_query_len: usize,
    pub b_g1_query: *const u64, pub b_g1_query_len: usize,
    pub b_g2_query: *const u64, pub b_g2_query_len: usize,
    pub l_query: *const u64, pub l_query_len: usize,
    pub h_query: *const u64, pub h_query_len: usize,
    pub window_bits: c_int,
}

extern "C" {
    pub fn cs_last_error() -> *const c_char;
    pub fn cs_ctx_create(device: c_int, stream: *mut c_void, out: *mut *mut cs_ctx) -> c_int;
    pub fn cs_ctx_destroy(ctx: *mut cs_ctx);
    pub fn cs_os_random(out: *mut u8, bytes: usize) -> c_int;
    // --- msm / fft (seam 1)
    pub fn cs_bases_upload(ctx: *mut cs_ctx, curve: c_int, group: c_int, pts: *const u64, n: usize,
                           window_bits: c_int, out: *mut *mut cs_bases) -> c_int;
    pub fn cs_bases_free(b: *mut cs_bases);
    pub fn cs_msm(ctx: *mut cs_ctx, b: *const cs_bases, offset: usize, scalars: *const u64, n: usize,
                  scalars_montgomery: c_int, out_affine: *mut u64, out_is_inf: *mut c_int) -> c_int;
    pub fn cs_domain_create(ctx: *mut cs_ctx, curve: c_int, log_n: c_uint, gen: *const u64,
                            out: *mut *mut cs_domain) -> c_int;
    pub fn cs_domain_free(d: *mut cs_domain);
    pub fn cs_ifft_in_to_out_host(ctx: *mut cs_ctx, d: *const cs_domain, data: *mut u64, batch: c_uint) -> c_int;
    pub fn cs_fft_out_to_in_host(ctx: *mut cs_ctx, d: *const cs_domain, data: *mut u64, batch: c_uint) -> c_int;
    // --- Groth16 (seams 2, 3)
    pub fn cs_groth16_pk_create(ctx: *mut cs_ctx, d: *const cs_groth16_key_desc, out: *mut *mut cs_groth16_pk) -> c_int;
    pub fn cs_groth16_pk_free(pk: *mut cs_groth16_pk);
    pub fn cs_groth16_domain_size(pk: *const cs_groth16_pk) -> usize;
    pub fn cs_groth16_witness_map(ctx: *mut cs_ctx, pk: *mut cs_groth16_pk, kind: c_int, party: c_int,
                                  public_inputs: *const u64, witness: *const u64, mask1: *const u64,
                                  mask2: *const u64, h_out: *mut u64) -> c_int;
    pub fn cs_groth16_prove_plain(ctx: *mut cs_ctx, pk: *mut cs_groth16_pk, public_inputs: *const u64,
                                  witness: *const u64, r: *const u64, s: *const u64, out_a: *mut u64,
                                  out_b: *mut u64, out_c: *mut u64) -> c_int;
    // --- transport + Rep3 / Shamir parties inside the library
    pub fn cs_net_from_callbacks(id: c_int, n_parties: c_int, cb: *const cs_net_callbacks, out: *mut *mut cs_net) -> c_int;
    pub fn cs_net_free(net: *mut cs_net);
    pub fn cs_rep3_state_create(net: *mut cs_net, out: *mut *mut cs_rep3_state) -> c_int;
    pub fn cs_rep3_state_from_seeds(party: c_int, own: *const u8, pos_own: u64, prev: *const u8, pos_prev: u64,
                                    out: *mut *mut cs_rep3_state) -> c_int;
    pub fn cs_rep3_state_free(st: *mut cs_rep3_state);
    pub fn cs_groth16_rep3_prove(ctx: *mut cs_ctx, pk: *mut cs_groth16_pk, net0: *mut cs_net, net1: *mut cs_net,
                                 state: *mut cs_rep3_state, public_inputs: *const u64, h_witness_shares: *const u64,
                                 d_witness_shares: *const u64, out_a: *mut u64, out_b: *mut u64, out_c: *mut u64,
                                 out_rs: *mut u64) -> c_int;
    pub fn cs_groth16_shamir_prove(ctx: *mut cs_ctx, pk: *mut cs_groth16_pk, net0: *mut cs_net, net1: *mut cs_net,
                                   num_parties: c_int, threshold: c_int, public_inputs: *const u64,
                                   witness_shares: *const u64, out_a: *mut u64, out_b: *mut u64, out_c: *mut u64,
                                   out_rs: *mut u64) -> c_int;
    pub fn cs_groth16_prove_with_shamir_bridge(ctx: *mut cs_ctx, pk: *mut cs_groth16_pk, net0: *mut cs_net,
                                               net1: *mut cs_net, public_inputs: *const u64,
                                               witness_rep3_shares: *const u64, out_a: *mut u64, out_b: *mut u64,
                                               out_c: *mut u64, out_rs: *mut u64) -> c_int;
}

/// `Err(message)` for a non-zero return code.
pub fn check(rc: c_int) -> Result<(), String> {
    if rc == 0 {
        return Ok(());
    }
    let msg = unsafe { std::ffi::CStr::from_ptr(cs_last_error()) }.to_string_lossy().into_owned();
    Err(format!("cosnarks_gpu ({rc}): {msg}"))
}
