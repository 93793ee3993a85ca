//! Numbers crossing the boundary.  The device holds every carry entry as a two's complement integer of
//! `limbs` little-endian 64-bit words over ONE common positive denominator (include/relp_gpu.h); relp works
//! with normalised `RationalBig`s.  `relp-num` exposes `FromStr` / `Display` in the form `n/d`
//!, which is the only constructor this file relies on.
use std::str::FromStr;

use relp_num::RationalBig;

/// Magnitude of a two's complement limb array (little endian) and its sign.
fn magnitude(words: &[u64]) -> (bool, Vec<u64>) {
    let negative = words.last().map_or(false, |w| (*w as i64) < 0);
    let mut mag = words.to_vec();
    if negative {
        let mut carry = 1u64;
        for w in mag.iter_mut() {
            let (v, c) = (!*w).overflowing_add(carry);
            *w = v;
            carry = c as u64;
        }
    }
    (negative, mag)
}

/// Decimal string of a little-endian magnitude (schoolbook division by 10^19).
fn to_decimal(mut mag: Vec<u64>) -> String {
    const CHUNK: u64 = 10_000_000_000_000_000_000;
    let mut parts: Vec<u64> = Vec::new();
    while mag.iter().any(|w| *w != 0) {
        let mut rem: u128 = 0;
        for w in mag.iter_mut().rev() {
            let cur = (rem << 64) | *w as u128;
            *w = (cur / CHUNK as u128) as u64;
            rem = cur % CHUNK as u128;
        }
        parts.push(rem as u64);
    }
    match parts.pop() {
        None => "0".to_string(),
        Some(top) => {
            let mut s = top.to_string();
            for p in parts.iter().rev() {
                s.push_str(&format!("{p:019}"));
            }
            s
        }
    }
}

/// numerator (two's complement limbs) / denominator (positive limbs)  ->  normalised `RationalBig`.
pub fn rational_from_limbs(numerator: &[u64], denominator: &[u64]) -> RationalBig {
    let (negative, mag) = magnitude(numerator);
    let text = format!("{}{}/{}", if negative { "-" } else { "" }, to_decimal(mag), to_decimal(denominator.to_vec()));
    RationalBig::from_str(&text).expect("n/d is relp-num's own Display format")
}

/// A vector of `count` numbers of `limbs` words each over one denominator; zeros are dropped, which is the
/// `SparseVector` convention of relp (data/linear_algebra/vector/sparse.rs:90-103).
pub fn sparse_from_limbs(words: &[u64], limbs: usize, count: usize, denominator: &[u64]) -> Vec<(usize, RationalBig)> {
    (0..count)
        .filter_map(|i| {
            let w = &words[i * limbs..(i + 1) * limbs];
            if w.iter().all(|x| *x == 0) { None } else { Some((i, rational_from_limbs(w, denominator))) }
        })
        .collect()
}

/// `(numerator, denominator)` of a rational given in relp-num's `n/d` (or `n`) Display form, as i128.
/// The integer prescale needs coefficients below 2^63 after scaling; providers with
/// larger coefficients are rejected by `GpuCarry::attach_provider`.
pub fn num_den<T: std::fmt::Display>(value: &T) -> (i128, i128) {
    let s = value.to_string();
    match s.split_once('/') {
        Some((n, d)) => (n.parse().expect("numerator"), d.parse().expect("denominator")),
        None => (s.parse().expect("integer"), 1),
    }
}

pub fn gcd(a: i128, b: i128) -> i128 { if b == 0 { a.abs() } else { gcd(b, a % b) } }
pub fn lcm(a: i128, b: i128) -> i128 { a / gcd(a, b) * b }
