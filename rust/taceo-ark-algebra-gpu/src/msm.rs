//! `msm_unchecked` / `msm_bigint`.  The base set of an MSM is a proving-key or SRS array that lives as long as the
//! prover, so its expanded table is uploaded once and cached by (pointer, length): `cs_bases_upload` is off the
//! per-proof path, `cs_msm` is the per-proof call.
use crate::{ctx, ids};
use ark_ec::short_weierstrass::{Affine, Projective, SWCurveConfig};
use ark_ec::AffineRepr;
use ark_ff::PrimeField;
use cosnarks_gpu_sys as sys;
use std::collections::HashMap;
use std::sync::Mutex;

struct Bases { ptr: *mut sys::cs_bases, start: usize, len: usize }
unsafe impl Send for Bases {}

static CACHE: Mutex<Option<HashMap<(usize, i32, i32), Vec<Bases>>>> = Mutex::new(None);

/// Packs `Affine<C>` (`{x, y, infinity}`, not `repr(C)`) into `x || y` limb arrays; infinity = all zero.
fn pack<C: SWCurveConfig>(points: &[Affine<C>]) -> Vec<u64> {
    let limbs = std::mem::size_of::<C::BaseField>() / 8;
    let mut out = vec![0u64; points.len() * 2 * limbs];
    for (i, p) in points.iter().enumerate() {
        if p.is_zero() { continue; }
        // BaseField = Fp<MontBackend<_, N>> (G1) or QuadExtField of it (G2: c0 || c1): plain Montgomery limbs
        let (x, y) = (p.x().unwrap(), p.y().unwrap());
        unsafe {
            std::ptr::copy_nonoverlapping(&x as *const _ as *const u64, out.as_mut_ptr().add(i * 2 * limbs), limbs);
            std::ptr::copy_nonoverlapping(&y as *const _ as *const u64, out.as_mut_ptr().add(i * 2 * limbs + limbs), limbs);
        }
    }
    out
}

/// Finds (or uploads) a resident base set that contains `points`; returns (handle, offset of points[0]).
fn bases_for<C: SWCurveConfig>(curve: i32, group: i32, points: &[Affine<C>]) -> (*mut sys::cs_bases, usize) {
    let key = (std::mem::size_of::<Affine<C>>(), curve, group);
    let addr = points.as_ptr() as usize;
    let mut guard = CACHE.lock().unwrap();
    let sets = guard.get_or_insert_with(HashMap::new).entry(key).or_default();
    let sz = std::mem::size_of::<Affine<C>>();
    for b in sets.iter() {
        // a sub-slice of an uploaded array (calculate_coeff slices query[1 + pub..], groth16.rs:190-200)
        if addr >= b.start && addr + points.len() * sz <= b.start + b.len * sz && (addr - b.start) % sz == 0 {
            return (b.ptr, (addr - b.start) / sz);
        }
    }
    let packed = pack(points);
    let mut h = std::ptr::null_mut();
    sys::check(unsafe { sys::cs_bases_upload(ctx(), curve, group, packed.as_ptr(), points.len(), 0, &mut h) })
        .expect("cs_bases_upload");
    sets.push(Bases { ptr: h, start: addr, len: points.len() });
    (h, 0)
}

fn run<C: SWCurveConfig>(points: &[Affine<C>], scalars: *const u64, n: usize, montgomery: i32) -> Option<Projective<C>> {
    let (curve, group) = ids::<C>()?;
    if n == 0 { return Some(Projective::<C>::default()); }
    let (bases, offset) = bases_for(curve, group, points);
    let limbs = std::mem::size_of::<C::BaseField>() / 8;
    let mut out = vec![0u64; 2 * limbs];
    let mut inf = 0i32;
    sys::check(unsafe { sys::cs_msm(ctx(), bases, offset, scalars, n, montgomery, out.as_mut_ptr(), &mut inf) })
        .expect("cs_msm");
    if inf != 0 { return Some(Projective::<C>::default()); }
    let (mut x, mut y) = (C::BaseField::default(), C::BaseField::default());
    unsafe {
        std::ptr::copy_nonoverlapping(out.as_ptr(), &mut x as *mut _ as *mut u64, limbs);
        std::ptr::copy_nonoverlapping(out.as_ptr().add(limbs), &mut y as *mut _ as *mut u64, limbs);
    }
    Some(Affine::<C>::new_unchecked(x, y).into())
}

/// `msm_unchecked(points, scalars)`: lengths may differ, the slices are chopped to the shorter (honk_curve.rs:33-34).
pub fn msm_unchecked<C: SWCurveConfig + 'static>(points: &[Affine<C>], scalars: &[C::ScalarField]) -> Projective<C> {
    let n = points.len().min(scalars.len());
    // &[Fr] is already [u64; 4] Montgomery limbs
    run(&points[..n], scalars.as_ptr().cast(), n, 1)
        .unwrap_or_else(|| taceo_ark_algebra_cpu::msm::msm_unchecked(points, scalars))
}

/// `msm_bigint(points, bigints)`: canonical (non-Montgomery) scalars (pointshare.rs:211-219).
pub fn msm_bigint<C: SWCurveConfig + 'static>(
    points: &[Affine<C>],
    bigints: &[<C::ScalarField as PrimeField>::BigInt],
) -> Projective<C> {
    let n = points.len().min(bigints.len());
    run(&points[..n], bigints.as_ptr().cast(), n, 0)
        .unwrap_or_else(|| taceo_ark_algebra_cpu::msm::msm_bigint(points, bigints))
}
