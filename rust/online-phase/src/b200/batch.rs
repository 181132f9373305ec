//! The whole-vector operator surface: `AuthenticatedScalarBatch<C>` is to a `Vec<AuthenticatedScalarResult<C>>` what a device
//! plane is to a `Vec<Scalar<C>>` — the same protocol, one result id and one kernel launch per gate instead of n.
//!
//! `batch_mul` below registers THREE operations where `AuthenticatedScalarResult::batch_mul` (authenticated_scalar.rs:848-879)
//! registers nine batch gates and a network op over 13n + 2 result ids:
//!
//! ```text
//!   mask       gate     (x, y, a, b)                    -> d_mine || e_mine          arkmpc_fr_beaver_mask
//!   exchange   network  d_mine || e_mine  <->  peer's    (NetworkPayload::ScalarBatch, unchanged wire type)
//!   recombine  gate     (mine, peer, a, b, c)           -> [x * y]                   arkmpc_fr_beaver_recombine
//! ```
//!
//! The closures capture an `Arc<Gates<C>>` and are `Send + Sync`; they run on whichever thread the executor uses
//! (fabric.rs:428, multi_threaded/executor.rs:208-217): the library serialises calls per context.
use std::sync::Arc;

use ark_ec::CurveGroup;
use itertools::Itertools;

use super::carrier::{ScalarBatchValue, ShareBatchValue};
use super::{B200Context, DeviceScalarBatch, DeviceShareBatch, Gates};
use sha3::{Digest, Sha3_256};

use crate::algebra::{AuthenticatedScalarResult, Scalar, ScalarResult, ScalarShare};
use crate::fabric::{MpcFabric, ResultHandle, ResultValue};
use crate::network::NetworkPayload;
use crate::{ResultId, PARTY0};

/// A vector of secret-shared, authenticated scalars held on the device under one result id
#[derive(Clone)]
pub struct AuthenticatedScalarBatch<C: CurveGroup> {
    /// Number of elements
    pub n: usize,
    /// The shares (resolves to a `DeviceShareBatch`)
    pub shares: ResultHandle<C, ShareBatchValue<C>>,
    gates: Arc<Gates<C>>,
}

/// A vector of public scalars held on the device under one result id
#[derive(Clone)]
pub struct ScalarBatch<C: CurveGroup> {
    /// Number of elements
    pub n: usize,
    /// The values (resolves to a `DeviceScalarBatch`)
    pub values: ResultHandle<C, ScalarBatchValue<C>>,
    gates: Arc<Gates<C>>,
}

/// The result of `open_authenticated`: the opened vector and one MAC-check flag for the whole batch, exactly as every
/// element of the reference's result vector shares one `commitment_check` (authenticated_scalar.rs:345-352)
#[derive(Clone)]
pub struct AuthenticatedScalarBatchOpenResult<C: CurveGroup> {
    /// The opened values
    pub value: ScalarBatch<C>,
    /// `Scalar::from(1)` iff the MAC check passed
    pub mac_check: ScalarResult<C>,
}

/// Create the per-fabric gate set: one native context on `device`, bound to this party's MAC key share
pub fn gates_for_fabric<C: CurveGroup>(fabric: &MpcFabric<C>, device: i32) -> Arc<Gates<C>> {
    let ctx = B200Context::new(device).unwrap_or_else(|e| panic!("no usable B200 context: {e:?}"));
    Arc::new(Gates::new(ctx, fabric.party_id(), fabric.mac_key()).unwrap_or_else(|e| panic!("{e:?}")))
}

impl<C: CurveGroup> AuthenticatedScalarBatch<C> {
    fn fabric(&self) -> &MpcFabric<C> {
        self.shares.fabric()
    }

    fn id(&self) -> ResultId {
        self.shares.id()
    }

    /// Gather a per-element vector (the reference's representation) into one device batch: ONE gate with n arguments
    pub fn from_elements(gates: &Arc<Gates<C>>, values: &[AuthenticatedScalarResult<C>]) -> Self {
        assert!(!values.is_empty(), "cannot gather an empty batch");
        let fabric = values[0].fabric().clone();
        let ids = values.iter().map(|v| v.id()).collect_vec();
        let g = gates.clone();
        let shares = fabric.new_gate_op(ids, move |args| {
            let host: Vec<ScalarShare<C>> = args.map(|a| a.into()).collect();
            ResultValue::from(DeviceShareBatch::upload(g.ctx(), &host))
        });
        Self { n: values.len(), shares, gates: gates.clone() }
    }

    /// Allocate host shares (e.g. a batch of Beaver triples from the `PreprocessingPhase`) as one device batch: one upload
    pub fn from_host_shares(fabric: &MpcFabric<C>, gates: &Arc<Gates<C>>, host: Vec<ScalarShare<C>>) -> Self {
        let n = host.len();
        let g = gates.clone();
        let shares = fabric.new_gate_op(vec![], move |_args| ResultValue::from(DeviceShareBatch::upload(g.ctx(), &host)));
        Self { n, shares, gates: gates.clone() }
    }

    /// Scatter back into the reference's per-element representation: ONE batch gate of output arity n
    pub fn to_elements(&self) -> Vec<AuthenticatedScalarResult<C>> {
        let n = self.n;
        // `AuthenticatedScalarResult<C>` is `ResultHandle<C, ScalarShare<C>>` (authenticated_scalar.rs:35)
        self.fabric().new_batch_gate_op(vec![self.id()], n, move |mut args| {
            let batch: ShareBatchValue<C> = args.next().unwrap().into();
            batch.download().into_iter().map(ResultValue::ScalarShare).collect_vec()
        })
    }

    /// The next n Beaver triples as three device batches (fabric.rs:894-915 without the 3n per-element allocations)
    pub fn next_triple_batch(fabric: &MpcFabric<C>, gates: &Arc<Gates<C>>, n: usize) -> (Self, Self, Self) {
        let (a, b, c) = fabric.inner.offline_phase.lock().expect("beaver source poisoned").next_triplet_batch(n);
        (
            Self::from_host_shares(fabric, gates, a),
            Self::from_host_shares(fabric, gates, b),
            Self::from_host_shares(fabric, gates, c),
        )
    }

    /// Element-wise authenticated Beaver multiplication (authenticated_scalar.rs:848-879)
    pub fn batch_mul(x: &Self, y: &Self) -> Self {
        assert_eq!(x.n, y.n, "Cannot compute batch mul on vectors of unequal length"); // :852
        let n = x.n;
        let fabric = x.fabric().clone();
        let gates = x.gates.clone();
        let (a, b, c) = Self::next_triple_batch(&fabric, &gates, n);

        // mask: d_mine || e_mine, the share components of [x - a] and [y - b] (:863-867; open_batch sends shares only, :141-145)
        let g = gates.clone();
        let de_mine: ResultHandle<C, ScalarBatchValue<C>> = fabric.new_gate_op(vec![x.id(), y.id(), a.id(), b.id()], move |mut args| {
            let x: ShareBatchValue<C> = args.next().unwrap().into();
            let y: ShareBatchValue<C> = args.next().unwrap().into();
            let a: ShareBatchValue<C> = args.next().unwrap().into();
            let b: ShareBatchValue<C> = args.next().unwrap().into();
            ResultValue::from(g.beaver_mask(&x, &y, &a, &b))
        });

        // exchange: the existing ScalarBatch payload, 2n scalars each way (party 0 sends first, fabric.rs:751-765)
        let de_peer = exchange_device_batch(&fabric, &gates, &de_mine);

        // recombine: open-add + de + d[b] + e[a] + [c] with the MAC update, one kernel (:161-171, :871-878)
        let g = gates.clone();
        let shares = fabric.new_gate_op(vec![de_mine.id(), de_peer.id(), a.id(), b.id(), c.id()], move |mut args| {
            let mine: ScalarBatchValue<C> = args.next().unwrap().into();
            let peer: ScalarBatchValue<C> = args.next().unwrap().into();
            let a: ShareBatchValue<C> = args.next().unwrap().into();
            let b: ShareBatchValue<C> = args.next().unwrap().into();
            let c: ShareBatchValue<C> = args.next().unwrap().into();
            ResultValue::from(g.beaver_recombine(&mine, &peer, &a, &b, &c))
        });
        Self { n, shares, gates }
    }

    /// Element-wise addition (:457-489)
    pub fn batch_add(a: &Self, b: &Self) -> Self {
        Self::linear(a, b, false)
    }

    /// Element-wise subtraction (:662-688)
    pub fn batch_sub(a: &Self, b: &Self) -> Self {
        Self::linear(a, b, true)
    }

    fn linear(a: &Self, b: &Self, sub: bool) -> Self {
        assert_eq!(a.n, b.n, "Cannot add batches of unequal length");
        let g = a.gates.clone();
        let shares = a.fabric().new_gate_op(vec![a.id(), b.id()], move |mut args| {
            let a: ShareBatchValue<C> = args.next().unwrap().into();
            let b: ShareBatchValue<C> = args.next().unwrap().into();
            ResultValue::from(g.share_add(&a, &b, sub))
        });
        Self { n: a.n, shares, gates: a.gates.clone() }
    }

    /// Add a public vector (:493-528)
    pub fn batch_add_public(a: &Self, v: &ScalarBatch<C>) -> Self {
        assert_eq!(a.n, v.n, "Cannot add batches of unequal length");
        let g = a.gates.clone();
        let shares = a.fabric().new_gate_op(vec![a.id(), v.values.id()], move |mut args| {
            let a: ShareBatchValue<C> = args.next().unwrap().into();
            let v: ScalarBatchValue<C> = args.next().unwrap().into();
            ResultValue::from(g.share_add_public(&a, &v))
        });
        Self { n: a.n, shares, gates: a.gates.clone() }
    }

    /// Multiply by a public vector (:883-916)
    pub fn batch_mul_public(a: &Self, v: &ScalarBatch<C>) -> Self {
        assert_eq!(a.n, v.n, "Cannot multiply batches of unequal length");
        let g = a.gates.clone();
        let shares = a.fabric().new_gate_op(vec![a.id(), v.values.id()], move |mut args| {
            let a: ShareBatchValue<C> = args.next().unwrap().into();
            let v: ScalarBatchValue<C> = args.next().unwrap().into();
            ResultValue::from(g.share_mul_public(&a, &v))
        });
        Self { n: a.n, shares, gates: a.gates.clone() }
    }

    /// Sum of the batch as a one-element batch (:563-576)
    pub fn sum(&self) -> Self {
        let g = self.gates.clone();
        let shares = self.fabric().new_gate_op(vec![self.id()], move |mut args| {
            let a: ShareBatchValue<C> = args.next().unwrap().into();
            ResultValue::from(g.share_sum(&a))
        });
        Self { n: 1, shares, gates: self.gates.clone() }
    }

    /// Open without checking MACs (:129-172): exchange the share components, add
    pub fn open(&self) -> ScalarBatch<C> {
        let fabric = self.fabric().clone();
        let gates = self.gates.clone();
        let g = gates.clone();
        // the share plane alone, as its own device batch (what `share.share()` extracts per element, :141-145)
        let mine: ResultHandle<C, ScalarBatchValue<C>> = fabric.new_gate_op(vec![self.id()], move |mut args| {
            let a: ShareBatchValue<C> = args.next().unwrap().into();
            let host = DeviceScalarBatch::<C> { n: a.n, plane: copy_plane(g.ctx(), &a), _c: Default::default() };
            ResultValue::from(host)
        });
        let peer = exchange_device_batch(&fabric, &gates, &mine);
        let g = gates.clone();
        let values = fabric.new_gate_op(vec![mine.id(), peer.id()], move |mut args| {
            let m: ScalarBatchValue<C> = args.next().unwrap().into();
            let p: ScalarBatchValue<C> = args.next().unwrap().into();
            ResultValue::from(g.scalar_add(&m, &p))
        });
        ScalarBatch { n: self.n, values, gates }
    }

    /// Open and check the MACs (:278-354): the MAC-check vector is computed on the device, its canonical bytes feed the
    /// host SHA3 commitment (commitment.rs:63-89), and a failed check still surfaces as `MpcError::AuthenticationError`
    /// through the usual `mac_check` flag (:379-383)
    pub fn open_authenticated(&self) -> AuthenticatedScalarBatchOpenResult<C> {
        let fabric = self.fabric().clone();
        let gates = self.gates.clone();
        let opened = self.open();

        // mac_key * opened - mac, one kernel (:299-311)
        let g = gates.clone();
        let checks: ResultHandle<C, ScalarBatchValue<C>> = fabric.new_gate_op(vec![opened.values.id(), self.id()], move |mut args| {
            let o: ScalarBatchValue<C> = args.next().unwrap().into();
            let s: ShareBatchValue<C> = args.next().unwrap().into();
            ResultValue::from(g.mac_check(&o, &s))
        });

        // commit to the check values: H(check_0 || ... || check_{n-1} || blinder) over canonical big-endian bytes, SHA3-256, the
        // digest reduced mod p (commitment.rs:63-89).  The device produces the byte string; the hash stays on the host.
        let blinder = Scalar::<C>::random(&mut rand::thread_rng());
        let g = gates.clone();
        let my_comm: ScalarResult<C> = fabric.new_gate_op(vec![checks.id()], move |mut args| {
            let c: ScalarBatchValue<C> = args.next().unwrap().into();
            ResultValue::Scalar(commit_bytes::<C>(g.to_bytes_be(&c), blinder))
        });
        let peer_comm = fabric.exchange_value(my_comm);
        let peer_checks = exchange_device_batch(&fabric, &gates, &checks);
        let peer_blinder = fabric.exchange_value(fabric.allocate_scalar(blinder));

        let g = gates.clone();
        let mac_check: ScalarResult<C> =
            fabric.new_gate_op(vec![peer_checks.id(), peer_blinder.id(), peer_comm.id(), checks.id()], move |mut args| {
                let peer: ScalarBatchValue<C> = args.next().unwrap().into();
                let peer_blinder: Scalar<C> = args.next().unwrap().into();
                let peer_comm: Scalar<C> = args.next().unwrap().into();
                let mine: ScalarBatchValue<C> = args.next().unwrap().into();
                let opens = commit_bytes::<C>(g.to_bytes_be(&peer), peer_blinder) == peer_comm; // commitment.rs:33-47
                let ok = opens && g.sum_is_zero(&mine, &peer); // :201-220
                ResultValue::Scalar(Scalar::from(ok))
            });
        AuthenticatedScalarBatchOpenResult { value: opened, mac_check }
    }
}

impl<C: CurveGroup> ScalarBatch<C> {
    /// Upload public host values as one device batch
    pub fn from_host(fabric: &MpcFabric<C>, gates: &Arc<Gates<C>>, host: Vec<Scalar<C>>) -> Self {
        let n = host.len();
        let g = gates.clone();
        let values = fabric.new_gate_op(vec![], move |_args| ResultValue::from(DeviceScalarBatch::upload(g.ctx(), &host)));
        Self { n, values, gates: gates.clone() }
    }

    /// Scatter into the reference's per-element `ScalarResult`s
    pub fn to_elements(&self) -> Vec<ScalarResult<C>> {
        self.values.fabric().new_batch_gate_op(vec![self.values.id()], self.n, move |mut args| {
            let batch: ScalarBatchValue<C> = args.next().unwrap().into();
            batch.download().into_iter().map(ResultValue::Scalar).collect_vec()
        })
    }
}

/// Send a device batch as `NetworkPayload::ScalarBatch` and receive the peer's as a device batch.  Received values are checked
/// to be canonical residues before any gate may consume them (the closure path gets this from arkworks' deserialisation).
fn exchange_device_batch<C: CurveGroup>(
    fabric: &MpcFabric<C>,
    gates: &Arc<Gates<C>>,
    mine: &ResultHandle<C, ScalarBatchValue<C>>,
) -> ResultHandle<C, ScalarBatchValue<C>> {
    let send = |fabric: &MpcFabric<C>| {
        let _sent: ResultHandle<C, Vec<Scalar<C>>> = fabric.new_network_op(vec![mine.id()], |mut args| {
            let batch: ScalarBatchValue<C> = args.next().unwrap().into();
            NetworkPayload::ScalarBatch(batch.download())
        });
    };
    let received: ResultHandle<C, Vec<Scalar<C>>> = if fabric.party_id() == PARTY0 {
        send(fabric);
        fabric.receive_value()
    } else {
        let r = fabric.receive_value();
        send(fabric);
        r
    };
    let g = gates.clone();
    fabric.new_gate_op(vec![received.id()], move |mut args| {
        let host: Vec<Scalar<C>> = args.next().unwrap().into();
        let dev = DeviceScalarBatch::upload(g.ctx(), &host);
        assert!(g.validate(&dev), "peer sent a non-canonical field element");
        ResultValue::from(dev)
    })
}

/// `HashCommitment`'s digest (commitment.rs:33-47, :63-89) over an already serialised value string
fn commit_bytes<C: CurveGroup>(mut bytes: Vec<u8>, blinder: Scalar<C>) -> Scalar<C> {
    bytes.append(&mut blinder.to_bytes_be());
    let mut hasher = Sha3_256::new();
    hasher.update(bytes);
    Scalar::from_be_bytes_mod_order(hasher.finalize().as_slice())
}

/// Device-to-device copy of the share plane of a batch
fn copy_plane<C: CurveGroup>(ctx: &Arc<B200Context>, a: &DeviceShareBatch<C>) -> super::DevBuf {
    use std::os::raw::c_void;
    let out = super::DevBuf::new(ctx, a.n * super::SCALAR_BYTES);
    if a.n > 0 {
        #[allow(unsafe_code)]
        let rc = unsafe {
            super::sys::arkmpc_memcpy_d2d(ctx.raw(), out.as_mut_ptr() as *mut c_void, a.share.as_ptr() as *const c_void, a.n * super::SCALAR_BYTES)
        };
        ctx.check(rc, "arkmpc_memcpy_d2d");
    }
    out
}

/// The reference signature, served by the device path: gather, multiply, scatter.  Under `cfg(feature = "b200")` the body of
/// `AuthenticatedScalarResult::batch_mul` (authenticated_scalar.rs:848-879) becomes
/// `crate::b200::batch::batch_mul_elements(fabric.b200_gates(), a, b)` when `field_id::<C>()` is `Ok` and the batch is large
/// enough to be worth one upload (a few thousand gates); otherwise the arkworks closures run as before.
pub fn batch_mul_elements<C: CurveGroup>(
    gates: &Arc<Gates<C>>,
    a: &[AuthenticatedScalarResult<C>],
    b: &[AuthenticatedScalarResult<C>],
) -> Vec<AuthenticatedScalarResult<C>> {
    assert_eq!(a.len(), b.len(), "Cannot compute batch mul on vectors of unequal length");
    if a.is_empty() {
        return vec![]; // :854-856
    }
    let x = AuthenticatedScalarBatch::from_elements(gates, a);
    let y = AuthenticatedScalarBatch::from_elements(gates, b);
    AuthenticatedScalarBatch::batch_mul(&x, &y).to_elements()
}
