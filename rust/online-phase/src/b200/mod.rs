//! Blackwell back end for the batched online-phase gates: safe wrappers over `libarkmpc_b200` (sys.rs) and the
//! device-resident batch carriers that let ONE `ResultId` denote a whole vector.
//!
//! Drop-in: copy `rust/online-phase/{build.rs, src/b200}` into the reference crate, add `pub mod b200;` behind
//! `#[cfg(feature = "b200")]` in `lib.rs`, the `b200` feature and `build = "build.rs"` in `Cargo.toml`, and the two
//! `ResultValue` variants described in `carrier.rs`.  Nothing else in the crate changes: the per-element operator surface
//! (`AuthenticatedScalarResult::batch_mul(&[..], &[..]) -> Vec<..>`) keeps its signatures, and `batch.rs` adds the
//! whole-vector surface on top of the same fabric (`new_gate_op`, `new_network_op`, `receive_value`).
//!
//! Memory contract (include/arkmpc_b200.h): `Scalar<C>` wraps `C::ScalarField = Fp256<MontBackend<_, 4>>`, i.e. four
//! little-endian `u64` limbs holding the canonical Montgomery residue (algebra/scalar/scalar.rs:46), and `ScalarShare<C>` is
//! `{share, mac}` (algebra/scalar/share.rs:32-37).  A `&[Scalar<C>]` / `&[ScalarShare<C>]` therefore crosses the boundary as a
//! `*const u64` with no conversion.  Only BN254 Fr and Curve25519 Fr are native; any other curve keeps the arkworks closures.
#![allow(unsafe_code)]

pub mod batch;
pub mod carrier;
pub mod sys;

use std::ffi::CStr;
use std::marker::PhantomData;
use std::os::raw::{c_int, c_void};
use std::ptr;
use std::sync::Arc;

use ark_ec::CurveGroup;
use ark_ff::{BigInteger, PrimeField};

use crate::algebra::{Scalar, ScalarShare};
use crate::PartyId;

/// Bytes of one field element in memory
pub const SCALAR_BYTES: usize = 32;

/// Why the native path cannot serve a curve or a call
#[derive(Debug, Clone, PartialEq, Eq)]
pub enum B200Error {
    /// `C::ScalarField` is neither BN254 Fr nor Curve25519 Fr
    UnsupportedField,
    /// The library reported a failure: (status, detail)
    Native(i32, String),
}

/// The native field id of `C::ScalarField`, chosen from its modulus
pub fn field_id<C: CurveGroup>() -> Result<c_int, B200Error> {
    // little-endian u64 limbs of the two supported moduli
    const BN254_FR: [u64; 4] = [0x43e1f593f0000001, 0x2833e84879b97091, 0xb85045b68181585d, 0x30644e72e131a029];
    const C25519_FR: [u64; 4] = [0x5812631a5cf5d3ed, 0x14def9dea2f79cd6, 0x0000000000000000, 0x1000000000000000];
    let m = <C::ScalarField as PrimeField>::MODULUS;
    let limbs: Vec<u64> = m.to_bytes_le().chunks(8).map(|c| u64::from_le_bytes(c.try_into().unwrap())).collect();
    if std::mem::size_of::<C::ScalarField>() != SCALAR_BYTES {
        return Err(B200Error::UnsupportedField);
    }
    if limbs[..] == BN254_FR {
        Ok(sys::ARKMPC_BN254_FR)
    } else if limbs[..] == C25519_FR {
        Ok(sys::ARKMPC_CURVE25519_FR)
    } else {
        Err(B200Error::UnsupportedField)
    }
}

/// One native context: bound to one device, owned by one fabric.  The library serialises concurrent calls, so the context may
/// be shared by the executor thread and by rayon workers (multi_threaded/executor.rs:208-217).
pub struct B200Context {
    raw: *mut sys::ArkmpcCtx,
}

// The library takes a per-context lock in every entry point and keeps no thread-local state but the last error string
// (same pattern as mp-spdz-rs/src/ffi.rs:170-184).
unsafe impl Send for B200Context {}
unsafe impl Sync for B200Context {}

impl B200Context {
    /// Create a context on `device`
    pub fn new(device: i32) -> Result<Arc<Self>, B200Error> {
        let mut raw = ptr::null_mut();
        let rc = unsafe { sys::arkmpc_ctx_create(device, &mut raw) };
        if rc != sys::ARKMPC_OK {
            return Err(B200Error::Native(rc, "arkmpc_ctx_create".to_string()));
        }
        Ok(Arc::new(Self { raw }))
    }

    /// The raw handle, for the `sys` calls in this module tree
    pub(crate) fn raw(&self) -> *mut sys::ArkmpcCtx {
        self.raw
    }

    /// Detail of the calling thread's last failure
    pub fn last_error(&self) -> String {
        let p = unsafe { sys::arkmpc_last_error(self.raw) };
        if p.is_null() {
            return String::new();
        }
        unsafe { CStr::from_ptr(p) }.to_string_lossy().into_owned()
    }

    /// The reference panics on misuse (authenticated_scalar.rs:852, fabric/result.rs:127-133); native failures follow suit
    pub(crate) fn check(&self, rc: c_int, what: &str) {
        if rc != sys::ARKMPC_OK {
            panic!("{what}: arkmpc status {rc}: {}", self.last_error());
        }
    }

    /// Block until everything enqueued on the context has finished
    pub fn sync(&self) {
        let rc = unsafe { sys::arkmpc_ctx_sync(self.raw) };
        self.check(rc, "arkmpc_ctx_sync");
    }

    /// The next Beaver kernel does not depend on the one launched just before it (see the header)
    pub fn hint_independent(&self) {
        unsafe { sys::arkmpc_ctx_hint_independent(self.raw) };
    }

    /// Return the device blocks `DevBuf` drops have parked in the library's cache to the driver (synchronises the device)
    pub fn mem_trim(&self) {
        let rc = unsafe { sys::arkmpc_mem_trim(self.raw) };
        self.check(rc, "arkmpc_mem_trim");
    }

    /// Bytes currently parked in the cache of free device blocks
    pub fn mem_cached_bytes(&self) -> usize {
        let mut bytes: usize = 0;
        let rc = unsafe { sys::arkmpc_mem_cached_bytes(self.raw, &mut bytes) };
        self.check(rc, "arkmpc_mem_cached_bytes");
        bytes
    }
}

impl Drop for B200Context {
    fn drop(&mut self) {
        unsafe { sys::arkmpc_ctx_destroy(self.raw) };
    }
}

/// A device allocation, freed on drop.  `arkmpc_free` parks the block in a per-device cache and orders its reuse on the device
/// after the work already submitted to every live context's stream, so dropping a buffer never synchronises the host — the
/// allocation pattern of one `ResultValue` per gate result stays cheap (include/arkmpc_b200.h, "memory").
pub struct DevBuf {
    ctx: Arc<B200Context>,
    ptr: *mut c_void,
    bytes: usize,
}

unsafe impl Send for DevBuf {}
unsafe impl Sync for DevBuf {}

impl DevBuf {
    /// Allocate `bytes` bytes (zero bytes allocate nothing)
    pub fn new(ctx: &Arc<B200Context>, bytes: usize) -> Self {
        let mut ptr = ptr::null_mut();
        let rc = unsafe { sys::arkmpc_malloc(ctx.raw(), bytes, &mut ptr) };
        ctx.check(rc, "arkmpc_malloc");
        Self { ctx: ctx.clone(), ptr, bytes }
    }

    /// Device pointer viewed as limbs
    pub fn as_ptr(&self) -> *const u64 {
        self.ptr as *const u64
    }

    /// Mutable device pointer viewed as limbs
    pub fn as_mut_ptr(&self) -> *mut u64 {
        self.ptr as *mut u64
    }

    /// Device pointer `elems` field elements into the buffer
    pub fn offset(&self, elems: usize) -> *mut u64 {
        assert!(elems * SCALAR_BYTES <= self.bytes, "offset past the end of a device buffer");
        unsafe { (self.ptr as *mut u8).add(elems * SCALAR_BYTES) as *mut u64 }
    }

    /// Size in bytes
    pub fn len_bytes(&self) -> usize {
        self.bytes
    }
}

impl Drop for DevBuf {
    fn drop(&mut self) {
        if !self.ptr.is_null() {
            unsafe { sys::arkmpc_free(self.ctx.raw(), self.ptr) };
        }
    }
}

/// A vector of public scalars on the device: one plane of `n` elements
pub struct DeviceScalarBatch<C: CurveGroup> {
    /// Number of scalars
    pub n: usize,
    /// The plane
    pub plane: DevBuf,
    _c: PhantomData<C>,
}

/// A vector of `ScalarShare`s on the device: a share plane and a mac plane of `n` elements each
pub struct DeviceShareBatch<C: CurveGroup> {
    /// Number of shares
    pub n: usize,
    /// Share components
    pub share: DevBuf,
    /// MAC components
    pub mac: DevBuf,
    _c: PhantomData<C>,
}

impl<C: CurveGroup> DeviceScalarBatch<C> {
    /// Uninitialised plane of `n` scalars
    pub fn alloc(ctx: &Arc<B200Context>, n: usize) -> Self {
        Self { n, plane: DevBuf::new(ctx, n * SCALAR_BYTES), _c: PhantomData }
    }

    /// Upload host scalars (their in-memory Montgomery image is what the device computes on)
    pub fn upload(ctx: &Arc<B200Context>, values: &[Scalar<C>]) -> Self {
        let out = Self::alloc(ctx, values.len());
        if !values.is_empty() {
            let rc = unsafe {
                sys::arkmpc_memcpy_h2d(ctx.raw(), out.plane.ptr, values.as_ptr() as *const c_void, values.len() * SCALAR_BYTES)
            };
            ctx.check(rc, "arkmpc_memcpy_h2d");
            // the copy is asynchronous on the context stream and `values` is only borrowed: finish it before returning
            ctx.sync();
        }
        out
    }

    /// Download into host scalars
    pub fn download(&self) -> Vec<Scalar<C>> {
        let ctx = &self.plane.ctx;
        let mut out = vec![Scalar::<C>::zero(); self.n];
        if self.n > 0 {
            let rc = unsafe {
                sys::arkmpc_memcpy_d2h(ctx.raw(), out.as_mut_ptr() as *mut c_void, self.plane.ptr, self.n * SCALAR_BYTES)
            };
            ctx.check(rc, "arkmpc_memcpy_d2h");
            ctx.sync();
        }
        out
    }
}

impl<C: CurveGroup> DeviceShareBatch<C> {
    /// Uninitialised planes for `n` shares
    pub fn alloc(ctx: &Arc<B200Context>, n: usize) -> Self {
        Self { n, share: DevBuf::new(ctx, n * SCALAR_BYTES), mac: DevBuf::new(ctx, n * SCALAR_BYTES), _c: PhantomData }
    }

    /// Upload the AoS image of `Vec<ScalarShare<C>>` and split it into planes on the device (arkmpc_share_unzip)
    pub fn upload(ctx: &Arc<B200Context>, shares: &[ScalarShare<C>]) -> Self {
        let n = shares.len();
        let out = Self::alloc(ctx, n);
        if n > 0 {
            let aos = DevBuf::new(ctx, 2 * n * SCALAR_BYTES);
            let rc = unsafe { sys::arkmpc_memcpy_h2d(ctx.raw(), aos.ptr, shares.as_ptr() as *const c_void, 2 * n * SCALAR_BYTES) };
            ctx.check(rc, "arkmpc_memcpy_h2d");
            let rc = unsafe { sys::arkmpc_share_unzip(ctx.raw(), n, aos.as_ptr(), out.share.as_mut_ptr(), out.mac.as_mut_ptr()) };
            ctx.check(rc, "arkmpc_share_unzip");
            ctx.sync(); // `shares` is only borrowed, and `aos` is freed on return
        }
        out
    }

    /// Download as `Vec<ScalarShare<C>>`
    pub fn download(&self) -> Vec<ScalarShare<C>> {
        let ctx = &self.share.ctx;
        let mut out = vec![ScalarShare::<C>::default(); self.n];
        if self.n > 0 {
            let aos = DevBuf::new(ctx, 2 * self.n * SCALAR_BYTES);
            let rc = unsafe { sys::arkmpc_share_zip(ctx.raw(), self.n, self.share.as_ptr(), self.mac.as_ptr(), aos.as_mut_ptr()) };
            ctx.check(rc, "arkmpc_share_zip");
            let rc = unsafe { sys::arkmpc_memcpy_d2h(ctx.raw(), out.as_mut_ptr() as *mut c_void, aos.ptr, 2 * self.n * SCALAR_BYTES) };
            ctx.check(rc, "arkmpc_memcpy_d2h");
            ctx.sync();
        }
        out
    }
}

/// The gate kernels, one method per fused gate.  `key` is the party's MAC key share (`MpcFabric::mac_key`).
pub struct Gates<C: CurveGroup> {
    ctx: Arc<B200Context>,
    field: c_int,
    party: PartyId,
    key: Scalar<C>,
}

impl<C: CurveGroup> Gates<C> {
    /// Bind the gates of one party to a context; fails for curves the library does not implement
    pub fn new(ctx: Arc<B200Context>, party: PartyId, key: Scalar<C>) -> Result<Self, B200Error> {
        Ok(Self { ctx, field: field_id::<C>()?, party, key })
    }

    /// The context
    pub fn ctx(&self) -> &Arc<B200Context> {
        &self.ctx
    }

    fn key_ptr(&self) -> *const u64 {
        &self.key as *const Scalar<C> as *const u64
    }

    /// Beaver phase 1 (authenticated_scalar.rs:863-867 as far as open_batch consumes it): d_mine || e_mine, 2n scalars
    pub fn beaver_mask(
        &self,
        x: &DeviceShareBatch<C>,
        y: &DeviceShareBatch<C>,
        a: &DeviceShareBatch<C>,
        b: &DeviceShareBatch<C>,
    ) -> DeviceScalarBatch<C> {
        let n = x.n;
        assert!(y.n == n && a.n == n && b.n == n, "batch_mul operands must have equal length"); // :852
        let de = DeviceScalarBatch::<C>::alloc(&self.ctx, 2 * n);
        let rc = unsafe {
            sys::arkmpc_fr_beaver_mask(
                self.ctx.raw(), self.field, n, x.share.as_ptr(), y.share.as_ptr(), a.share.as_ptr(), b.share.as_ptr(),
                de.plane.as_mut_ptr(), de.plane.offset(n),
            )
        };
        self.ctx.check(rc, "arkmpc_fr_beaver_mask");
        de
    }

    /// Beaver phase 2: open-add of d, e fused with `de + d[b] + e[a] + [c]` and the MAC update (:161-171, :871-878)
    pub fn beaver_recombine(
        &self,
        de_mine: &DeviceScalarBatch<C>,
        de_peer: &DeviceScalarBatch<C>,
        a: &DeviceShareBatch<C>,
        b: &DeviceShareBatch<C>,
        c: &DeviceShareBatch<C>,
    ) -> DeviceShareBatch<C> {
        let n = a.n;
        assert!(de_mine.n == 2 * n && de_peer.n == 2 * n && b.n == n && c.n == n, "recombine operands must match");
        let out = DeviceShareBatch::<C>::alloc(&self.ctx, n);
        let rc = unsafe {
            sys::arkmpc_fr_beaver_recombine(
                self.ctx.raw(), self.field, self.party as c_int, self.key_ptr(), n,
                de_mine.plane.as_ptr(), de_mine.plane.offset(n), de_peer.plane.as_ptr(), de_peer.plane.offset(n),
                a.share.as_ptr(), a.mac.as_ptr(), b.share.as_ptr(), b.mac.as_ptr(), c.share.as_ptr(), c.mac.as_ptr(),
                out.share.as_mut_ptr(), out.mac.as_mut_ptr(), ptr::null_mut(), ptr::null_mut(),
            )
        };
        self.ctx.check(rc, "arkmpc_fr_beaver_recombine");
        out
    }

    /// `batch_add` / `batch_sub` on shares (:457-489, :662-688)
    pub fn share_add(&self, a: &DeviceShareBatch<C>, b: &DeviceShareBatch<C>, sub: bool) -> DeviceShareBatch<C> {
        assert_eq!(a.n, b.n, "batch_add operands must have equal length");
        let out = DeviceShareBatch::<C>::alloc(&self.ctx, a.n);
        let f = if sub { sys::arkmpc_fr_share_sub } else { sys::arkmpc_fr_share_add };
        let rc = unsafe {
            f(self.ctx.raw(), self.field, a.n, a.share.as_ptr(), a.mac.as_ptr(), b.share.as_ptr(), b.mac.as_ptr(),
              out.share.as_mut_ptr(), out.mac.as_mut_ptr())
        };
        self.ctx.check(rc, "arkmpc_fr_share_add/sub");
        out
    }

    /// `batch_add_public` (:493-528; share.rs:74-77: party 0 adds the value, both add mac_key * value to the MAC)
    pub fn share_add_public(&self, a: &DeviceShareBatch<C>, v: &DeviceScalarBatch<C>) -> DeviceShareBatch<C> {
        assert_eq!(a.n, v.n, "batch_add_public operands must have equal length");
        let out = DeviceShareBatch::<C>::alloc(&self.ctx, a.n);
        let rc = unsafe {
            sys::arkmpc_fr_share_add_public(
                self.ctx.raw(), self.field, self.party as c_int, self.key_ptr(), a.n, a.share.as_ptr(), a.mac.as_ptr(),
                v.plane.as_ptr(), out.share.as_mut_ptr(), out.mac.as_mut_ptr(),
            )
        };
        self.ctx.check(rc, "arkmpc_fr_share_add_public");
        out
    }

    /// `batch_mul_public` (:883-916)
    pub fn share_mul_public(&self, a: &DeviceShareBatch<C>, v: &DeviceScalarBatch<C>) -> DeviceShareBatch<C> {
        assert_eq!(a.n, v.n, "batch_mul_public operands must have equal length");
        let out = DeviceShareBatch::<C>::alloc(&self.ctx, a.n);
        let rc = unsafe {
            sys::arkmpc_fr_share_mul_public(
                self.ctx.raw(), self.field, a.n, a.share.as_ptr(), a.mac.as_ptr(), v.plane.as_ptr(),
                out.share.as_mut_ptr(), out.mac.as_mut_ptr(),
            )
        };
        self.ctx.check(rc, "arkmpc_fr_share_mul_public");
        out
    }

    /// Element-wise sum of two public vectors (the open-add of `open_batch`, :166-168)
    pub fn scalar_add(&self, a: &DeviceScalarBatch<C>, b: &DeviceScalarBatch<C>) -> DeviceScalarBatch<C> {
        assert_eq!(a.n, b.n, "operands must have equal length");
        let out = DeviceScalarBatch::<C>::alloc(&self.ctx, a.n);
        let rc = unsafe { sys::arkmpc_fr_add(self.ctx.raw(), self.field, a.n, a.plane.as_ptr(), b.plane.as_ptr(), out.plane.as_mut_ptr()) };
        self.ctx.check(rc, "arkmpc_fr_add");
        out
    }

    /// The MAC-check vector of `open_authenticated_batch`: mac_key * opened - mac (:299-311)
    pub fn mac_check(&self, opened: &DeviceScalarBatch<C>, shares: &DeviceShareBatch<C>) -> DeviceScalarBatch<C> {
        assert_eq!(opened.n, shares.n, "operands must have equal length");
        let out = DeviceScalarBatch::<C>::alloc(&self.ctx, opened.n);
        let rc = unsafe {
            sys::arkmpc_fr_mac_check(self.ctx.raw(), self.field, self.key_ptr(), opened.n, opened.plane.as_ptr(), shares.mac.as_ptr(), out.plane.as_mut_ptr())
        };
        self.ctx.check(rc, "arkmpc_fr_mac_check");
        out
    }

    /// `mine[i] + peer[i] == 0` for every i (:217-219).  Synchronises.
    pub fn sum_is_zero(&self, mine: &DeviceScalarBatch<C>, peer: &DeviceScalarBatch<C>) -> bool {
        assert_eq!(mine.n, peer.n, "operands must have equal length");
        let mut flag: c_int = 0;
        let rc = unsafe { sys::arkmpc_fr_sum_is_zero(self.ctx.raw(), self.field, mine.n, mine.plane.as_ptr(), peer.plane.as_ptr(), &mut flag) };
        self.ctx.check(rc, "arkmpc_fr_sum_is_zero");
        flag != 0
    }

    /// Values that arrived from the peer must be canonical residues before any gate consumes them (what arkworks'
    /// deserialisation enforces for the closure path, scalar.rs:187-202).  Synchronises.
    pub fn validate(&self, v: &DeviceScalarBatch<C>) -> bool {
        let mut flag: c_int = 0;
        let rc = unsafe { sys::arkmpc_fr_validate(self.ctx.raw(), self.field, v.n, v.plane.as_ptr(), &mut flag) };
        self.ctx.check(rc, "arkmpc_fr_validate");
        flag != 0
    }

    /// `Sum` of a share vector (:563-576, share.rs:104-111) as a one-element batch
    pub fn share_sum(&self, a: &DeviceShareBatch<C>) -> DeviceShareBatch<C> {
        let out = DeviceShareBatch::<C>::alloc(&self.ctx, 1);
        let rc = unsafe {
            sys::arkmpc_fr_share_sum(self.ctx.raw(), self.field, a.n, a.share.as_ptr(), a.mac.as_ptr(), out.share.as_mut_ptr(), out.mac.as_mut_ptr())
        };
        self.ctx.check(rc, "arkmpc_fr_share_sum");
        out
    }

    /// Canonical big-endian bytes of every element, the input of the hash commitment (scalar.rs:118-127, commitment.rs:63-89)
    pub fn to_bytes_be(&self, v: &DeviceScalarBatch<C>) -> Vec<u8> {
        let bytes = DevBuf::new(&self.ctx, v.n * SCALAR_BYTES);
        let mut out = vec![0u8; v.n * SCALAR_BYTES];
        if v.n > 0 {
            let rc = unsafe { sys::arkmpc_fr_to_bytes_be(self.ctx.raw(), self.field, v.n, v.plane.as_ptr(), bytes.ptr as *mut u8) };
            self.ctx.check(rc, "arkmpc_fr_to_bytes_be");
            let rc = unsafe { sys::arkmpc_memcpy_d2h(self.ctx.raw(), out.as_mut_ptr() as *mut c_void, bytes.ptr, out.len()) };
            self.ctx.check(rc, "arkmpc_memcpy_d2h");
            self.ctx.sync();
        }
        out
    }
}
