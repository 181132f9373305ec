//! The device-batch result carrier: one `ResultId` denotes a whole vector that lives on the GPU.
//!
//! Two variants are added to `enum ResultValue<C>` (fabric/result.rs:47-64):
//!
//! ```ignore
//!     /// A batch of public scalars resident on the device (one plane)
//!     #[cfg(feature = "b200")]
//!     DeviceScalarBatch(std::sync::Arc<crate::b200::DeviceScalarBatch<C>>),
//!     /// A batch of scalar shares resident on the device (share plane + mac plane)
//!     #[cfg(feature = "b200")]
//!     DeviceShareBatch(std::sync::Arc<crate::b200::DeviceShareBatch<C>>),
//! ```
//!
//! plus their arms in `impl Debug for ResultValue` (`f.debug_tuple("DeviceScalarBatch").field(&b.n).finish()`).  `ResultValue`
//! is `Clone` and the executor clones every argument of every gate (single_threaded.rs:339), which is why the payload is an
//! `Arc`: cloning a 2^20-element batch costs one reference count instead of 64 MiB.  Results are written once
//! (result_buffer.rs:3-8), so sharing is sound.  The casts below follow the existing ones (fabric/result.rs:127-133): a
//! mismatch panics.  Device batches never travel as such: the network ops in batch.rs turn them into the existing
//! `NetworkPayload::ScalarBatch` (network.rs:45-60), so neither transport changes.
use std::sync::Arc;

use ark_ec::CurveGroup;

use super::{DeviceScalarBatch, DeviceShareBatch};
use crate::fabric::ResultValue;

/// A handle-side alias: what a `ResultHandle<C, _>` of a device batch resolves to
pub type ScalarBatchValue<C> = Arc<DeviceScalarBatch<C>>;
/// As above for share batches
pub type ShareBatchValue<C> = Arc<DeviceShareBatch<C>>;

impl<C: CurveGroup> From<ResultValue<C>> for Arc<DeviceScalarBatch<C>> {
    fn from(value: ResultValue<C>) -> Self {
        match value {
            ResultValue::DeviceScalarBatch(batch) => batch,
            _ => panic!("Cannot cast {:?} to a device scalar batch", value),
        }
    }
}

impl<C: CurveGroup> From<Arc<DeviceScalarBatch<C>>> for ResultValue<C> {
    fn from(value: Arc<DeviceScalarBatch<C>>) -> Self {
        ResultValue::DeviceScalarBatch(value)
    }
}

impl<C: CurveGroup> From<DeviceScalarBatch<C>> for ResultValue<C> {
    fn from(value: DeviceScalarBatch<C>) -> Self {
        ResultValue::DeviceScalarBatch(Arc::new(value))
    }
}

impl<C: CurveGroup> From<ResultValue<C>> for Arc<DeviceShareBatch<C>> {
    fn from(value: ResultValue<C>) -> Self {
        match value {
            ResultValue::DeviceShareBatch(batch) => batch,
            _ => panic!("Cannot cast {:?} to a device share batch", value),
        }
    }
}

impl<C: CurveGroup> From<Arc<DeviceShareBatch<C>>> for ResultValue<C> {
    fn from(value: Arc<DeviceShareBatch<C>>) -> Self {
        ResultValue::DeviceShareBatch(value)
    }
}

impl<C: CurveGroup> From<DeviceShareBatch<C>> for ResultValue<C> {
    fn from(value: DeviceShareBatch<C>) -> Self {
        ResultValue::DeviceShareBatch(Arc::new(value))
    }
}
