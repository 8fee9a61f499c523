// maxdecoy-sys/src/lib.rs -- Rust view of include/maxdecoy.h.
// Shipped as source: the build image has no rustc/cargo, so this file is not compiled here; the ctypes mirror
// max-decoy_b200/maxdecoy/_abi.py exercises the same ABI in the tests.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

pub enum md_ctx {}

#[repr(C)] pub struct md_config { pub device: i32, pub n_threads: u32 }
#[repr(C)] pub struct md_modification {
    pub accession: [c_char; 24], pub name: [c_char; 40],
    pub position: u8, pub is_fix: u8, pub amino_acid: u8, pub _pad: [u8; 5], pub mono_mass: i64,
}
#[repr(C)] pub struct md_digest_params { pub max_missed_cleavages: u32, pub min_len: u32, pub max_len: u32 }
#[repr(C)] pub struct md_peptide_table {
    pub n: u64, pub seq_bytes: u64, pub n_assoc: u64,
    pub seq: *mut u8, pub seq_off: *mut u64, pub missed_cleavages: *mut u8, pub weight: *mut i64,
    pub counts: *mut i16, pub assoc_off: *mut u64, pub assoc_protein: *mut u32,
}
#[repr(C)] pub struct md_index_stats { pub n_peptides: u64, pub seq_bytes: u64, pub device_bytes: u64, pub min_key: i64, pub max_key: i64 }
#[repr(C)] pub struct md_precursor { pub mass: i64, pub lo: i64, pub hi: i64, pub charge: u32, pub spectrum_id: u32 }
#[repr(C)] pub struct md_candidate_table {
    pub n_spectra: u32, pub n: u64, pub off: *mut u64, pub peptide_id: *mut u64, pub var_mask: *mut u64, pub mod_weight: *mut i64,
}
#[repr(C)] pub struct md_decoy_table {
    pub n_spectra: u32, pub n: u64, pub seq_bytes: u64, pub off: *mut u64, pub seq: *mut u8, pub seq_off: *mut u64,
    pub var_mask: *mut u64, pub weight: *mut i64, pub mod_weight: *mut i64, pub attempt: *mut u32,
}
#[repr(C)] pub struct md_spectra {
    pub n: u32, pub precursor_mz: *const f64, pub charge: *const u8, pub spectrum_id: *const u32,
    pub peak_off: *const u64, pub peak_mz: *const f64, pub peak_intensity: *const f32,
}
#[repr(C)] pub struct md_search_params {
    pub lower_ppm: i64, pub upper_ppm: i64, pub abs_lower_uda: i64, pub abs_upper_uda: i64, pub fragment_tolerance: f64,
    pub n_decoys: u32, pub decoy_mode: i32, pub seed: u64, pub top_k: u32, pub min_peaks: u32,
    pub max_fragment_charge: u32, pub keep_decoys: u32,
}
#[repr(C)] pub struct md_psm {
    pub spectrum_id: u32, pub rank: u16, pub is_decoy: u8, pub charge: u8, pub candidate: u64, pub var_mask: u64,
    pub mod_weight: i64, pub raw_score: i64, pub score: f32, pub n_targets: u32, pub n_decoys: u32, pub _pad: u32,
}   // 56 bytes
#[repr(C)] pub struct md_identify_stats {
    pub n_spectra: u64, pub n_targets: u64, pub n_decoys: u64, pub n_less_decoys: u64, pub n_kernel_launches: u64,
    pub ms_lookup: f64, pub ms_decoys: f64, pub ms_score: f64, pub ms_total: f64, pub ms_kernel_score: f64,
    pub ms_kernel_decoy: f64, pub n_attempts: u64, pub n_pairs: u64, pub score_bytes: u64,
    pub ms_score_prepare: f64, pub n_score_left: u64, pub score_pipelined: u32, pub _pad: u32,
}

extern "C" {
    pub fn md_create(cfg: *const md_config, out: *mut *mut md_ctx) -> c_int;
    pub fn md_destroy(ctx: *mut md_ctx);
    pub fn md_last_error(ctx: *const md_ctx) -> *const c_char;
    pub fn md_backend_name() -> *const c_char;
    pub fn md_residue_mass(one_letter_code: u8) -> i64;
    pub fn md_sequence_weight(seq: *const u8, len: u32) -> i64;
    pub fn md_precursor_window(mz: f64, charge: u32, lower_ppm: i64, upper_ppm: i64, p: *mut i64, lo: *mut i64, hi: *mut i64) -> c_int;
    pub fn md_set_modifications(ctx: *mut md_ctx, mods: *const md_modification, n: u32, max_variable_mods: u32) -> c_int;
    pub fn md_substitution_map(ctx: *mut md_ctx, out441: *mut i64) -> c_int;
    pub fn md_digest(ctx: *mut md_ctx, residues: *const u8, protein_offsets: *const u64, n_proteins: u32,
                     params: *const md_digest_params, n_peptides: *mut u64) -> c_int;
    pub fn md_peptides_export(ctx: *mut md_ctx, out: *mut md_peptide_table) -> c_int;
    pub fn md_peptide_table_free(t: *mut md_peptide_table);
    pub fn md_index_build(ctx: *mut md_ctx) -> c_int;
    pub fn md_index_stats_get(ctx: *mut md_ctx, out: *mut md_index_stats) -> c_int;
    pub fn md_index_export(ctx: *mut md_ctx, begin: u64, count: u64, peptide_id: *mut u64, key: *mut i64) -> c_int;
    pub fn md_window_search(ctx: *mut md_ctx, lo: *const i64, hi: *const i64, n: u32, begin: *mut u64, end: *mut u64) -> c_int;
    pub fn md_candidates(ctx: *mut md_ctx, p: *const md_precursor, n: u32, out: *mut md_candidate_table) -> c_int;
    pub fn md_candidate_table_free(t: *mut md_candidate_table);
    pub fn md_decoy_store_set(ctx: *mut md_ctx, seq: *const u8, seq_off: *const u64, n: u64) -> c_int;
    pub fn md_set_variable_mode(ctx: *mut md_ctx, mode: c_int) -> c_int;
    pub fn md_generate_decoys(ctx: *mut md_ctx, p: *const md_precursor, n_spectra: u32, n_per_spectrum: u32, mode: c_int,
                              seed: u64, out: *mut md_decoy_table) -> c_int;
    pub fn md_decoy_table_free(t: *mut md_decoy_table);
    pub fn md_identify(ctx: *mut md_ctx, spectra: *const md_spectra, params: *const md_search_params, psms: *mut md_psm,
                       stats: *mut md_identify_stats, all_scores: *mut *mut i64, all_off: *mut *mut u64) -> c_int;
    pub fn md_identify_device(ctx: *mut md_ctx, spectra_dev: *const md_spectra, params: *const md_search_params, psms_dev: *mut md_psm,
                              stats: *mut md_identify_stats) -> c_int;
    pub fn md_comm_unique_id(id: *mut u8) -> c_int;                       // MD_COMM_ID_BYTES = 128
    pub fn md_comm_init(ctx: *mut md_ctx, rank: i32, nranks: i32, id: *const u8) -> c_int;
    pub fn md_gather_psms(ctx: *mut md_ctx, local: *const md_psm, rows_per_rank: u64, all: *mut md_psm) -> c_int;
    pub fn md_comm_destroy(ctx: *mut md_ctx) -> c_int;
    pub fn md_sync(ctx: *mut md_ctx) -> c_int;
    pub fn md_stream_handle(ctx: *mut md_ctx) -> *mut c_void;
    pub fn md_last_decoys_export(ctx: *mut md_ctx, out: *mut md_decoy_table) -> c_int;
    pub fn md_free(p: *mut c_void);
}
