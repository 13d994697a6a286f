//! Seam 1: `[patch.crates-io] taceo-ark-algebra`.  Same public items as taceo-ark-algebra 0.1.0 as the reference
//! uses them (call sites: co-groth16/src/mpc/{plain.rs:73, rep3.rs:131, shamir.rs:118}, groth16.rs:194,
//! groth16/reduction.rs:93,141-175,249,270-328, mpc-core/src/protocols/rep3/pointshare.rs:217-218,
//! co-noir-common/src/honk_curve.rs:82).  Everything that is not BN254 / BLS12-381 x {G1, G2} x
//! {Fr, Rep3 share, Shamir share} falls through to the CPU crate.
pub mod fft;
pub mod msm;

use cosnarks_gpu_sys as sys;
use std::any::TypeId;
use std::cell::RefCell;

thread_local! {
    /// one cs_ctx per rayon worker: the reference calls msm / fft from up to five workers at once (groth16.rs:227)
    static CTX: RefCell<Option<*mut sys::cs_ctx>> = RefCell::new(None);
}

pub(crate) fn ctx() -> *mut sys::cs_ctx {
    CTX.with(|c| {
        *c.borrow_mut().get_or_insert_with(|| {
            let dev: i32 = std::env::var("COSNARKS_GPU_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
            let mut p = std::ptr::null_mut();
            sys::check(unsafe { sys::cs_ctx_create(dev, std::ptr::null_mut(), &mut p) }).expect("cs_ctx_create");
            p
        })
    })
}

/// (curve id, group id) of a short-Weierstrass config the library has kernels for.
pub(crate) fn ids<C: 'static>() -> Option<(i32, i32)> {
    let t = TypeId::of::<C>();
    if t == TypeId::of::<ark_bn254::g1::Config>() { Some((sys::CS_BN254, sys::CS_G1)) }
    else if t == TypeId::of::<ark_bn254::g2::Config>() { Some((sys::CS_BN254, sys::CS_G2)) }
    else if t == TypeId::of::<ark_bls12_381::g1::Config>() { Some((sys::CS_BLS12_381, sys::CS_G1)) }
    else if t == TypeId::of::<ark_bls12_381::g2::Config>() { Some((sys::CS_BLS12_381, sys::CS_G2)) }
    else { None }
}
