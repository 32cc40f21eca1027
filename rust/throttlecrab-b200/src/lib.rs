//! `throttlecrab-b200`: throttlecrab's `Store` trait and `rate_limit` call over libgcra_b200.so.
//! NOT compiled in this repository's environment (no rustc/cargo in the image): source for maintainers.
//! The C ABI it binds is include/gcra_b200.h; tests of the same ABI run through ctypes (tests/) and C++ (examples/).
use std::ffi::c_void;
use std::time::{Duration, SystemTime, UNIX_EPOCH};
use throttlecrab::{CellError, RateLimitResult, Store};

#[repr(C)] pub struct GcraConfig { capacity: u64, device: i32, store_kind: i32, p0: u64, p1: u64, p2: u64,
                                   created_ns: i64, max_batch: u32, flags: u32, hash_seed: [u64; 2] }
#[repr(C)] #[derive(Clone, Copy)] pub struct GcraRequest { pub key_hash: u64, pub max_burst: i64,
    pub count_per_period: i64, pub period: i64, pub quantity: i64, pub now_ns: i64 }
#[repr(C)] #[derive(Clone, Copy, Default)] pub struct GcraResult { pub remaining: i64, pub reset_after_ns: i64,
    pub retry_after_ns: i64, pub status: i32, pub allowed: u8, pad: [u8; 3] }
#[repr(C)] pub struct GcraEngine { _p: [u8; 0] }

extern "C" {
    fn gcra_create(cfg: *const GcraConfig, out: *mut *mut GcraEngine) -> i32;
    fn gcra_destroy(h: *mut GcraEngine);
    fn gcra_last_error(h: *mut GcraEngine) -> *const std::os::raw::c_char;
    fn gcra_hash_key(key: *const c_void, len: u64) -> u64;
    fn gcra_store_get(h: *mut GcraEngine, key: *const c_void, len: u64, now_ns: i64, value: *mut i64, found: *mut u8) -> i32;
    fn gcra_store_cas(h: *mut GcraEngine, key: *const c_void, len: u64, old: i64, new: i64, ttl_ns: u64, now_ns: i64, ok: *mut u8) -> i32;
    fn gcra_store_set_nx(h: *mut GcraEngine, key: *const c_void, len: u64, value: i64, ttl_ns: u64, now_ns: i64, ok: *mut u8) -> i32;
    fn gcra_rate_limit(h: *mut GcraEngine, key: *const c_void, len: u64, max_burst: i64, count: i64, period: i64,
                       quantity: i64, now_ns: i64, out: *mut GcraResult) -> i32;
    fn gcra_rate_limit_batch(h: *mut GcraEngine, n: u64, req: *const GcraRequest, res: *mut GcraResult) -> i32;
}

fn ns(t: SystemTime) -> i64 { t.duration_since(UNIX_EPOCH).map(|d| d.as_nanos() as i64).unwrap_or(-1) }

/// `impl Store`: the reference's own generic `RateLimiter<S: Store>` runs unchanged over the GPU table.
pub struct GpuStore { h: *mut GcraEngine }
unsafe impl Send for GpuStore {}

impl GpuStore {
    pub fn adaptive(capacity: usize) -> Result<Self, String> { Self::new(capacity, 2) }
    pub fn periodic(capacity: usize) -> Result<Self, String> { Self::new(capacity, 0) }
    fn new(capacity: usize, kind: i32) -> Result<Self, String> {
        let cfg = GcraConfig { capacity: capacity as u64, device: 0, store_kind: kind, p0: 0, p1: 0, p2: 0,
                               created_ns: ns(SystemTime::now()), max_batch: 0, flags: 8 /* GCRA_FLAG_RANDOM_SEED */,
                               hash_seed: [0, 0] };
        let mut h = std::ptr::null_mut();
        if unsafe { gcra_create(&cfg, &mut h) } != 0 { return Err("gcra_create failed (no CUDA device)".into()); }
        Ok(GpuStore { h })
    }
    fn err(&self) -> String { unsafe { std::ffi::CStr::from_ptr(gcra_last_error(self.h)) }.to_string_lossy().into() }
}
impl Drop for GpuStore { fn drop(&mut self) { unsafe { gcra_destroy(self.h) } } }

impl Store for GpuStore {
    fn get(&self, key: &str, now: SystemTime) -> Result<Option<i64>, String> {
        let (mut v, mut f) = (0i64, 0u8);
        if unsafe { gcra_store_get(self.h, key.as_ptr() as _, key.len() as u64, ns(now), &mut v, &mut f) } != 0 { return Err(self.err()); }
        Ok(if f != 0 { Some(v) } else { None })
    }
    fn compare_and_swap_with_ttl(&mut self, key: &str, old: i64, new: i64, ttl: Duration, now: SystemTime) -> Result<bool, String> {
        let mut ok = 0u8;
        if unsafe { gcra_store_cas(self.h, key.as_ptr() as _, key.len() as u64, old, new, ttl.as_nanos() as u64, ns(now), &mut ok) } != 0 { return Err(self.err()); }
        Ok(ok != 0)
    }
    fn set_if_not_exists_with_ttl(&mut self, key: &str, value: i64, ttl: Duration, now: SystemTime) -> Result<bool, String> {
        let mut ok = 0u8;
        if unsafe { gcra_store_set_nx(self.h, key.as_ptr() as _, key.len() as u64, value, ttl.as_nanos() as u64, ns(now), &mut ok) } != 0 { return Err(self.err()); }
        Ok(ok != 0)
    }
}

/// Same signature as `RateLimiter::rate_limit`, one fused call instead of get + CAS round trips,
/// and the batched form the actor should use.
pub struct GpuRateLimiter { store: GpuStore }
impl GpuRateLimiter {
    pub fn new(store: GpuStore) -> Self { Self { store } }
    pub fn rate_limit(&mut self, key: &str, max_burst: i64, count_per_period: i64, period: i64, quantity: i64,
                      now: SystemTime) -> Result<(bool, RateLimitResult), CellError> {
        let mut r = GcraResult::default();
        let st = unsafe { gcra_rate_limit(self.store.h, key.as_ptr() as _, key.len() as u64, max_burst,
                                          count_per_period, period, quantity, ns(now), &mut r) };
        match st {
            0 => Ok((r.allowed != 0, RateLimitResult { limit: max_burst, remaining: r.remaining,
                     reset_after: Duration::from_nanos(r.reset_after_ns as u64),
                     retry_after: Duration::from_nanos(r.retry_after_ns as u64) })),
            1 => Err(CellError::NegativeQuantity(quantity)),
            2 => Err(CellError::InvalidRateLimit),
            _ => Err(CellError::Internal(self.store.err())),
        }
    }
    /// `reqs[i].key_hash = hash_key(key_i)`; results as if applied in index order.
    pub fn rate_limit_batch(&mut self, reqs: &[GcraRequest], out: &mut [GcraResult]) -> Result<(), CellError> {
        assert_eq!(reqs.len(), out.len());
        if unsafe { gcra_rate_limit_batch(self.store.h, reqs.len() as u64, reqs.as_ptr(), out.as_mut_ptr()) } != 0 {
            return Err(CellError::Internal(self.store.err()));
        }
        Ok(())
    }
}
pub fn hash_key(key: &str) -> u64 { unsafe { gcra_hash_key(key.as_ptr() as _, key.len() as u64) } }
