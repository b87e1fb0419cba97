//! Prints known-answer vectors of the third-party crates and of the reference itself (see ../Cargo.toml).
//! Every value is deterministic; floats are emitted as IEEE-754 bit patterns (u32) so that nothing is lost in JSON.
use ndarray::Array2;
use rand::distributions::{Distribution, WeightedIndex};
use rand::rngs::StdRng;
use rand::seq::{IteratorRandom, SliceRandom};
use rand::{Rng, RngCore, SeedableRng};
use serde_json::{json, Value};
use std::path::PathBuf;
use vector_indexer::ivf_index::IvfIndex;
use vector_indexer::kmeans::{run_kmeans_mini_batch, run_kmeans_parallel};
use vector_indexer::utils::{calculate_max_iterations, calculate_num_clusters, euclidean_distance_squared};
use vector_indexer::vector_store::VectorStore;
use wide::{f32x4, f32x8};

fn bits(v: impl IntoIterator<Item = f32>) -> Vec<u32> {
    v.into_iter().map(|x| x.to_bits()).collect()
}
fn hex(b: &[u8]) -> String {
    b.iter().map(|x| format!("{:02x}", x)).collect()
}
/// tests/test_utils/mod.rs:10-16
fn ramp(n: usize, dim: usize) -> Array2<f32> {
    Array2::from_shape_vec((n, dim), (0..n * dim).map(|x| (x as f32 * 0.1) % 50.0).collect()).unwrap()
}

fn main() {
    let out_path = std::env::args().nth(1).expect("usage: vidx_golden <output.json>");
    let mut g = serde_json::Map::new();

    // ---- rand 0.8.5 / rand_chacha 0.3.1 / rand_core 0.6.4 --------------------------------------------------------
    let mut r = StdRng::seed_from_u64(42);
    g.insert("rng_u32_seed42".into(), json!((0..130).map(|_| r.next_u32()).collect::<Vec<u32>>()));
    let mut r = StdRng::seed_from_u64(42);
    g.insert("rng_u64_seed42".into(), json!((0..40).map(|_| r.next_u64().to_string()).collect::<Vec<String>>()));
    // next_u64 straddling the 64-word buffer: 63 u32 draws first
    let mut r = StdRng::seed_from_u64(7);
    for _ in 0..63 {
        r.next_u32();
    }
    g.insert("rng_u64_straddle_seed7".into(), json!((0..3).map(|_| r.next_u64().to_string()).collect::<Vec<String>>()));
    for seed in [0u64, 1309, 756, u64::MAX] {
        let mut r = StdRng::seed_from_u64(seed);
        g.insert(format!("rng_u32_seed{}", seed), json!((0..8).map(|_| r.next_u32()).collect::<Vec<u32>>()));
    }
    let mut r = StdRng::seed_from_u64(42);
    let mut v: Vec<usize> = (0..10).collect();
    v.shuffle(&mut r);
    g.insert("shuffle_10_seed42".into(), json!(v));
    let mut r = StdRng::seed_from_u64(7);
    let mut v: Vec<usize> = (0..1000).collect();
    v.shuffle(&mut r);
    g.insert("shuffle_1000_seed7".into(), json!(v));
    let mut gr = serde_json::Map::new();
    for n in [10usize, 1000, 50_000, 1_000_000, 3_000_000_000] {
        let mut r = StdRng::seed_from_u64(42);
        gr.insert(n.to_string(), json!((0..20).map(|_| r.gen_range(0..n)).collect::<Vec<usize>>()));
    }
    g.insert("gen_range_usize_seed42".into(), Value::Object(gr));
    let weights: Vec<f32> = (0..16).map(|i| ((i * 37) % 11 + 1) as f32 * 0.25).collect();
    let mut r = StdRng::seed_from_u64(42);
    let dist = WeightedIndex::new(&weights).unwrap();
    g.insert("weighted_index_16_seed42".into(), json!((0..20).map(|_| dist.sample(&mut r)).collect::<Vec<usize>>()));
    // weights of wildly different magnitude (the k-means++ weights are d^2 of squared distances)
    let weights2: Vec<f32> = (0..1000).map(|i| ((i * 7919) % 1013) as f32).map(|x| x * x * x * 1e-3).collect();
    let mut r = StdRng::seed_from_u64(756);
    let dist = WeightedIndex::new(&weights2).unwrap();
    g.insert("weighted_index_1000_seed756".into(), json!((0..20).map(|_| dist.sample(&mut r)).collect::<Vec<usize>>()));
    let mut r = StdRng::seed_from_u64(756);
    g.insert("choose_multiple_50_7_seed756".into(), json!((0..50usize).choose_multiple(&mut r, 7)));
    let mut r = StdRng::seed_from_u64(756);
    g.insert("choose_multiple_4096_64_seed756".into(), json!((0..4096usize).choose_multiple(&mut r, 64)));
    // tests/test_utils/mod.rs:245-252: create_deterministic_vectors(4, 8, 7)
    let mut r = StdRng::seed_from_u64(7);
    g.insert("gen_range_f32_m10_10_seed7".into(), json!(bits((0..32).map(|_| r.gen_range(-10.0f32..10.0)))));

    // ---- wide 0.7.33 ------------------------------------------------------------------------------------------------
    // lane values are exact squares chosen so that every candidate reduction tree gives a different sum
    let x8: [f32; 8] = [16.5, 2176.0, 0.203125, 6.75, 54.0, 1.8125, 19.5, 2944.0];
    let x4: [f32; 4] = [1.125, 5248.0, 2.6875, 1.0];
    let s8 = f32x8::from(x8) * f32x8::from(x8);
    let s4 = f32x4::from(x4) * f32x4::from(x4);
    g.insert("wide_f32x8_reduce_add".into(), json!(s8.reduce_add().to_bits()));
    g.insert("wide_f32x4_reduce_add".into(), json!(s4.reduce_add().to_bits()));
    g.insert("wide_target_features".into(), json!({"avx": cfg!(target_feature = "avx"), "sse3": cfg!(target_feature = "sse3"),
                                                   "sse2": cfg!(target_feature = "sse2")}));

    // ---- the reference itself ---------------------------------------------------------------------------------------
    g.insert("calculate_num_clusters".into(),
             json!([3usize, 5000, 9999, 10_000, 50_000, 99_999, 100_000, 1_000_000, 10_000_000].map(|n| (n, calculate_num_clusters(n), calculate_max_iterations(n)))));
    // (exactly representable inputs: no libm in the way)
    let a: Vec<f32> = (0..37).map(|i| ((i * 37) % 101) as f32 * 0.125 - 3.0).collect();
    let b: Vec<f32> = (0..37).map(|i| ((i * 53) % 97) as f32 * 0.0625).collect();
    g.insert("euclidean_distance_squared_37".into(), json!(euclidean_distance_squared(&a, &b).to_bits()));
    // mini-batch k-means on the ramp fixture (kmeans_tests.rs uses it throughout): brute-force and hierarchical assignment
    for (n, dim, k, iters) in [(5000usize, 32usize, 20usize, 50usize), (5000, 32, 150, 30), (300, 13, 7, 25)] {
        let data = ramp(n, dim);
        let (c, l) = run_kmeans_mini_batch(&data, k, iters, None, 42).unwrap();
        g.insert(format!("kmeans_mini_batch_ramp_{}x{}_k{}_it{}", n, dim, k, iters),
                 json!({"centroids": bits(c.iter().cloned()), "labels": l.to_vec()}));
    }
    let data = ramp(600, 8);
    let (c, l) = run_kmeans_parallel(&data, 5, 10, None, 42).unwrap();
    g.insert("kmeans_parallel_ramp_600x8_k5_it10".into(), json!({"centroids": bits(c.iter().cloned()), "labels": l.to_vec()}));
    // a tiny index on disk: 12 vectors of dimension 3 (-> the 8-byte padding of shards.rs:104-157), fixed timestamps
    let dir: PathBuf = std::env::temp_dir().join(format!("vidx_golden_{}", std::process::id()));
    let (index_dir, shards_dir) = (dir.join("index"), dir.join("shards"));
    std::fs::create_dir_all(&index_dir).unwrap();
    std::fs::create_dir_all(&shards_dir).unwrap();
    let records: Vec<(u64, Vec<f32>, u64)> =
        (0..12u64).map(|i| (1000 + i, (0..3).map(|j| i as f32 * 0.37 + j as f32).collect(), 1_700_000_000 + i)).collect();
    let store = VectorStore::new(records);
    let mut ix = IvfIndex::new(3);
    ix.fit_with_paths(&store, &shards_dir, 42);
    ix.save_to(&index_dir).unwrap();
    let mut files = serde_json::Map::new();
    files.insert("index.bin".into(), json!(hex(&std::fs::read(index_dir.join("index.bin")).unwrap())));
    let mut names: Vec<String> = std::fs::read_dir(&shards_dir).unwrap().map(|e| e.unwrap().file_name().into_string().unwrap()).collect();
    names.sort();
    for n in names {
        files.insert(n.clone(), json!(hex(&std::fs::read(shards_dir.join(&n)).unwrap())));
    }
    g.insert("index_12x3_seed42_files".into(), Value::Object(files));
    let _ = std::fs::remove_dir_all(&dir);

    std::fs::write(&out_path, serde_json::to_string_pretty(&Value::Object(g)).unwrap()).unwrap();
    eprintln!("wrote {}", out_path);
}
