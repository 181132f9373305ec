//! Reference-run generator for tests/golden ("outputs of the reference itself").
//!
//! Copied by `regen.sh` into `online-phase/examples/gen_golden.rs` of a checkout of renegade-fi/ark-mpc and run with
//! `cargo run --example gen_golden --features "test_helpers"`.  It replays the deterministic two-party case of the committed
//! fixture `tests/golden/scalar_golden.json["party_id_beaver_source"]` (the reference's own `PartyIDBeaverSource`,
//! offline_prep.rs:109-170: a = 2, b = 3, c = 6, MAC key share = party id, input mask 3) through the reference's
//! `execute_mock_mpc` (lib.rs:116-201; TestCurve = BN254) and prints, per party, the Montgomery memory image (4 LE u64 limbs, hex)
//! of every value that the fixture holds: the shares of x and y after `batch_share_scalar`, the triples, the `batch_mul` output
//! shares and the opened products.  `compare_reference.py` diffs that JSON with the fixture: equality PINS the oracle to the
//! reference for BN254 Fr.  (This file is test infrastructure; it has never been compiled in this repository's image, which has
//! no Rust toolchain.)
use ark_mpc::{
    algebra::{AuthenticatedScalarResult, Scalar, ScalarShare},
    test_helpers::{execute_mock_mpc, TestCurve},
    PARTY0,
};
use futures::future;

type S = Scalar<TestCurve>;

/// Montgomery limbs of a scalar, as the hex strings the fixture uses: `Fp256<MontBackend<_, 4>>` is `Fp(BigInt([u64; 4]), _)`
fn limbs(s: &S) -> Vec<String> {
    s.inner().0 .0.iter().map(|l| format!("{:016x}", l)).collect()
}

fn share(s: &ScalarShare<TestCurve>) -> Vec<Vec<String>> {
    vec![limbs(&s.share()), limbs(&s.mac())]
}

fn json_shares(v: &[ScalarShare<TestCurve>]) -> String {
    serde_json::to_string(&v.iter().map(share).collect::<Vec<_>>()).unwrap()
}

#[tokio::main]
async fn main() {
    // the plaintext inputs of the fixture: x = [5, 0, p - 1, 7], y = [9, 3, p - 1, 0], both shared by party 0
    let minus_one = -S::one();
    let xs = vec![S::from(5u8), S::from(0u8), minus_one, S::from(7u8)];
    let ys = vec![S::from(9u8), S::from(3u8), minus_one, S::from(0u8)];
    let n = xs.len();

    let (p0, p1) = execute_mock_mpc(|fabric| {
        let xs = xs.clone();
        let ys = ys.clone();
        async move {
            let x = fabric.batch_share_scalar(xs, PARTY0);
            let y = fabric.batch_share_scalar(ys, PARTY0);
            let prod = AuthenticatedScalarResult::batch_mul(&x, &y);
            let opened = AuthenticatedScalarResult::open_authenticated_batch(&prod);

            // `AuthenticatedScalarResult<C>` is `ResultHandle<C, ScalarShare<C>>`: awaiting it yields this party's share
            let x_sh = future::join_all(x.iter().cloned()).await;
            let y_sh = future::join_all(y.iter().cloned()).await;
            let prod_sh = future::join_all(prod.iter().cloned()).await;
            let opened: Vec<S> = future::join_all(opened).await.into_iter().map(|r| r.expect("MAC check failed")).collect();
            (x_sh, y_sh, prod_sh, opened)
        }
    })
    .await;

    // the triples PartyIDBeaverSource hands out are constants (offline_prep.rs:137-158); restate them so the fixture's
    // "triples" section is covered by the same file
    let triple = |party: u64| -> Vec<ScalarShare<TestCurve>> {
        let key = S::from(party);
        let (a, b, c) = if party == 0 { (1u64, 3u64, 2u64) } else { (1u64, 0u64, 4u64) };
        vec![
            ScalarShare::new(S::from(a), key * S::from(2u8)),
            ScalarShare::new(S::from(b), key * S::from(3u8)),
            ScalarShare::new(S::from(c), key * S::from(6u8)),
        ]
    };

    let mut out = String::from("{\"field\":\"bn254_fr\",\"generator\":\"renegade-fi/ark-mpc execute_mock_mpc (reference run)\",");
    out += &format!("\"n\":{},", n);
    out += &format!("\"x\":[{},{}],", json_shares(&p0.0), json_shares(&p1.0));
    out += &format!("\"y\":[{},{}],", json_shares(&p0.1), json_shares(&p1.1));
    out += &format!("\"triple\":[{},{}],", json_shares(&triple(0)), json_shares(&triple(1)));
    out += &format!("\"batch_mul\":[{},{}],", json_shares(&p0.2), json_shares(&p1.2));
    let opened: Vec<Vec<String>> = p0.3.iter().map(limbs).collect();
    assert_eq!(p0.3, p1.3, "the parties opened different values");
    out += &format!("\"opened\":{}}}", serde_json::to_string(&opened).unwrap());
    println!("{}", out);
}
