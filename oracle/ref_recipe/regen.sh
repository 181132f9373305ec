#!/bin/bash
# Pin the oracle to the REFERENCE: build renegade-fi/ark-mpc on a box that has its toolchain and network access to crates.io,
# run its own execute_mock_mpc on the deterministic PartyIDBeaverSource case, and compare the outputs with the committed fixture.
#
#   usage: oracle/ref_recipe/regen.sh /path/to/ark-mpc-checkout      (a copy is built; the checkout itself is not modified)
#
# Needs: rustup with nightly-2024-02-26 (online-phase/rust-toolchain), cargo able to fetch ark-ff / ark-ec / ark-bn254 0.4.
# Not runnable in this repository's build image (no cargo / rustc, no network): DESIGN.md §2 "parity unpinned".
set -euo pipefail
REF=${1:?path to a checkout of renegade-fi/ark-mpc}
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/../.." && pwd)
WORK=$(mktemp -d)
trap 'rm -rf "$WORK"' EXIT
cp -r "$REF" "$WORK/ark-mpc"
mkdir -p "$WORK/ark-mpc/online-phase/examples"
cp "$HERE/gen_golden.rs" "$WORK/ark-mpc/online-phase/examples/gen_golden.rs"
cd "$WORK/ark-mpc/online-phase"
# the example needs tokio's macros and serde_json, both already dependencies of the crate (Cargo.toml)
cargo +nightly-2024-02-26 run --release --example gen_golden --features "test_helpers" > "$ROOT/tests/golden/reference_party_id.json"
python "$HERE/compare_reference.py" "$ROOT/tests/golden/reference_party_id.json"
echo "tests/golden/reference_party_id.json written and equal to the committed fixture: commit it; tests/test_golden.py then reports 'pinned'."
