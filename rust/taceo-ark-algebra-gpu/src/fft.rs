//! `fft::{Domain, bit_reverse}` (co-groth16/src/groth16/reduction.rs:93,141-175,249,270-328).  `Domain` is generic
//! over the coefficient type `T: DomainCoeff<F>`; the GPU path covers T = Fr (batch 1), ShamirPrimeFieldShare<Fr>
//! (a transparent wrapper: batch 1) and Rep3PrimeFieldShare<Fr> (`{a, b}`: batch 2, transformed in place).
use crate::ctx;
use ark_ff::FftField;
use cosnarks_gpu_sys as sys;
use std::any::TypeId;
pub use taceo_ark_algebra_cpu::fft::{bit_reverse, DomainCoeff};

pub struct Domain<F: FftField> {
    size: usize,
    gpu: Option<*mut sys::cs_domain>,
    cpu: taceo_ark_algebra_cpu::fft::Domain<F>,
}
unsafe impl<F: FftField> Send for Domain<F> {}
unsafe impl<F: FftField> Sync for Domain<F> {}

fn curve_of<F: 'static>() -> Option<i32> {
    let t = TypeId::of::<F>();
    if t == TypeId::of::<ark_bn254::Fr>() { Some(sys::CS_BN254) }
    else if t == TypeId::of::<ark_bls12_381::Fr>() { Some(sys::CS_BLS12_381) }
    else { None }
}

/// components per element: size_of::<T>() / size_of::<F>() is 1 for Fr / Shamir shares and 2 for Rep3 shares
fn batch_of<T, F>() -> Option<u32> {
    match std::mem::size_of::<T>() / std::mem::size_of::<F>() {
        1 if std::mem::size_of::<T>() == std::mem::size_of::<F>() => Some(1),
        2 if std::mem::size_of::<T>() == 2 * std::mem::size_of::<F>() => Some(2),
        _ => None,
    }
}

impl<F: FftField + 'static> Domain<F> {
    fn build(size: usize, gen: Option<F>, cpu: taceo_ark_algebra_cpu::fft::Domain<F>) -> Self {
        let gpu = curve_of::<F>().and_then(|curve| {
            let mut h = std::ptr::null_mut();
            let g = gen.as_ref().map_or(std::ptr::null(), |g| g as *const F as *const u64);
            sys::check(unsafe { sys::cs_domain_create(ctx(), curve, size.trailing_zeros(), g, &mut h) }).ok().map(|_| h)
        });
        Self { size, gpu, cpu }
    }
    /// `Domain::with_group_gen(size, gen)` (reduction.rs:93): the snarkjs root of unity of groth16.rs:60-100
    pub fn with_group_gen(size: usize, group_gen: F) -> Option<Self> {
        let cpu = taceo_ark_algebra_cpu::fft::Domain::with_group_gen(size, group_gen)?;
        Some(Self::build(cpu.size(), Some(group_gen), cpu))
    }
    /// `Domain::new(min_size)` (reduction.rs:249): arkworks' generator
    pub fn new(min_size: usize) -> Option<Self> {
        let cpu = taceo_ark_algebra_cpu::fft::Domain::new(min_size)?;
        Some(Self::build(cpu.size(), None, cpu))
    }
    pub fn size(&self) -> usize { self.size }
    /// natural order in, bit-reversed order out, 1/n included
    pub fn ifft_in_to_out<T: DomainCoeff<F>>(&self, v: &mut [T]) {
        match (self.gpu, batch_of::<T, F>()) {
            (Some(d), Some(b)) if v.len() == self.size =>
                sys::check(unsafe { sys::cs_ifft_in_to_out_host(ctx(), d, v.as_mut_ptr().cast(), b) }).expect("ifft"),
            _ => self.cpu.ifft_in_to_out(v),
        }
    }
    /// bit-reversed order in, natural order out
    pub fn fft_out_to_in<T: DomainCoeff<F>>(&self, v: &mut [T]) {
        match (self.gpu, batch_of::<T, F>()) {
            (Some(d), Some(b)) if v.len() == self.size =>
                sys::check(unsafe { sys::cs_fft_out_to_in_host(ctx(), d, v.as_mut_ptr().cast(), b) }).expect("fft"),
            _ => self.cpu.fft_out_to_in(v),
        }
    }
}

impl<F: FftField> Drop for Domain<F> {
    fn drop(&mut self) { if let Some(d) = self.gpu { unsafe { sys::cs_domain_free(d) } } }
}
