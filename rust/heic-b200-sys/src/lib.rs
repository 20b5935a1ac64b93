//! SOURCE ONLY (never compiled here: no Rust toolchain in the build image).  `extern "C"` declarations for every entry
//! point of `include/heic_b200.h` (tests/test_capi_exports.py checks this list against the header), plus the safe
//! wrappers that replace the tile loop of `HeicDecoder::decode` (src/heic/decoder.rs:98-119) and
//! `SliceSegmentReader::read_data` (src/hevc/slice.rs:206).  Struct layouts are those of the header; the flat `fields`
//! arrays stand for its runs of 32-bit members (41 / 36 / 15 words: the offsets the header's structs have).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

// heic_status
pub const HEIC_OK: i32 = 0;
pub const HEIC_E_INVALID_ARG: i32 = -1;
pub const HEIC_E_UNSUPPORTED: i32 = -2;
pub const HEIC_E_BITSTREAM: i32 = -3;
pub const HEIC_E_NO_DEVICE: i32 = -4;
pub const HEIC_E_CUDA: i32 = -5;
pub const HEIC_E_NOMEM: i32 = -6;
// stage mask of heic_b200_batch_run_stages
pub const HEIC_STAGE_CABAC: u32 = 1;
pub const HEIC_STAGE_TRANSFORM: u32 = 2;
pub const HEIC_STAGE_INTRA: u32 = 4;
pub const HEIC_STAGE_DEBLOCK: u32 = 8;
pub const HEIC_STAGE_SAO: u32 = 16;
pub const HEIC_STAGE_COLOR: u32 = 32;
pub const HEIC_STAGE_ALL: u32 = 63;

pub const HEIC_MAX_ENTRY_POINTS: usize = 255;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct heic_scaling_list {
    pub list: [[[u8; 64]; 6]; 4],
    pub dc: [[u8; 6]; 2],
}

/// Mirrors `SequenceParameterSet` (src/hevc/grammar.rs:388-428); field order = include/heic_b200.h.
#[repr(C)]
#[derive(Clone, Copy)]
pub struct heic_sps {
    pub fields: [u32; 41],
    pub scaling_list: heic_scaling_list,
}

/// Mirrors `PictureParameterSet` (src/hevc/grammar.rs:511-548).
#[repr(C)]
#[derive(Clone, Copy)]
pub struct heic_pps {
    pub fields: [u32; 36],
    pub scaling_list: heic_scaling_list,
}

/// Mirrors `SliceSegmentHeader` (src/hevc/grammar.rs:551-572) + the un-escaped substream offsets.
#[repr(C)]
#[derive(Clone, Copy)]
pub struct heic_slice_header {
    pub fields: [u32; 15],
    pub entry_point_offset_minus1: [u32; HEIC_MAX_ENTRY_POINTS],
    pub slice_data_byte_offset: u32,
    pub substream_offset: [u32; HEIC_MAX_ENTRY_POINTS + 1],
}

#[repr(C)]
pub struct heic_tile_desc {
    pub rbsp: *const u8,
    pub rbsp_len: u32,
    pub nal_unit_type: u32,
    /// 0: `rbsp` is un-escaped; 1: raw NAL payload, offsets in raw byte counts (un-escaped on the GPU)
    pub escaped: u32,
    pub header: heic_slice_header,
}

#[repr(C)]
pub struct heic_image_desc {
    pub sps: heic_sps,
    pub pps: heic_pps,
    pub grid_rows: u32,
    pub grid_cols: u32,
    pub output_width: u32,
    pub output_height: u32,
    pub rotation_ccw_quarter_turns: u32,
    pub n_tiles: u32,
    pub tiles: *const heic_tile_desc,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct heic_tile_status {
    pub code: i32,
    pub bins_decoded: u32,
    pub ctus_decoded: u32,
    pub reserved: u32,
}

/// Container metadata the reference's integration test pins (tests/libheif_comparison.rs:102-111).
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct heic_file_info {
    pub primary_item_id: u32,
    pub ispe_width: u32,
    pub ispe_height: u32,
    pub rotation_ccw_quarter_turns: u32,
    pub rotated_width: u32,
    pub rotated_height: u32,
    pub luma_bits: u32,
    pub chroma_bits: u32,
    pub thumbnail_count: u32,
    pub item_count: u32,
    pub is_grid: u32,
}

/// Host copies of one tile's intermediate buffers (per-stage parity tests); any pointer may be null.
#[repr(C)]
pub struct heic_tile_dump {
    pub tu_map: *mut u32,
    pub tu_map_len: u32,
    pub coeff: [*mut i16; 3],
    pub coeff_len: [u32; 3],
    pub qp_map: *mut u8,
    pub qp_map_len: u32,
    pub sao: *mut u32,
    pub sao_len: u32,
    pub plane: [*mut u8; 3],
    pub plane_len: [u32; 3],
}

#[repr(C)]
pub struct heic_b200_ctx {
    _private: [u8; 0],
}
#[repr(C)]
pub struct heic_b200_file {
    _private: [u8; 0],
}
#[repr(C)]
pub struct heic_b200_job {
    _private: [u8; 0],
}
#[repr(C)]
pub struct heic_b200_batch {
    _private: [u8; 0],
}

extern "C" {
    pub fn heic_b200_abi_version() -> i32;
    pub fn heic_b200_last_error() -> *const c_char;
    pub fn heic_b200_create(device: i32, out_ctx: *mut *mut heic_b200_ctx) -> i32;
    pub fn heic_b200_destroy(ctx: *mut heic_b200_ctx);
    pub fn heic_b200_launch_count(ctx: *const heic_b200_ctx) -> u64;
    // ---- host-side restatement of the reference's parse layer ----
    pub fn heic_b200_remove_emulation_prevention(
        data: *const u8, len: usize, out: *mut u8, epb_pos: *mut u32, epb_cap: usize, n_epb: *mut usize,
    ) -> i64;
    pub fn heic_b200_rbsp_read_ue(data: *const u8, len: usize, bit_pos: *mut usize, out: *mut u32) -> i32;
    pub fn heic_b200_rbsp_read_se(data: *const u8, len: usize, bit_pos: *mut usize, out: *mut i32) -> i32;
    pub fn heic_b200_parse_sps(rbsp: *const u8, len: usize, out: *mut heic_sps) -> i32;
    pub fn heic_b200_parse_pps(rbsp: *const u8, len: usize, out: *mut heic_pps) -> i32;
    pub fn heic_b200_parse_slice_header(
        rbsp: *const u8, len: usize, nal_unit_type: u32, sps: *const heic_sps, pps: *const heic_pps,
        epb_pos: *const u32, n_epb: usize, out: *mut heic_slice_header,
    ) -> i32;
    pub fn heic_b200_parse_slice_header_raw(
        nal_payload: *const u8, len: usize, nal_unit_type: u32, sps: *const heic_sps, pps: *const heic_pps,
        out: *mut heic_slice_header,
    ) -> i32;
    // ---- HeifReader / HeicDecoder item walk ----
    pub fn heic_b200_file_open(data: *const u8, len: usize, out: *mut *mut heic_b200_file) -> i32;
    pub fn heic_b200_file_close(f: *mut heic_b200_file);
    pub fn heic_b200_file_primary_image(f: *const heic_b200_file) -> *const heic_image_desc;
    pub fn heic_b200_file_aux_image_count(f: *const heic_b200_file) -> u32;
    pub fn heic_b200_file_aux_image(f: *const heic_b200_file, i: u32) -> *const heic_image_desc;
    pub fn heic_b200_file_primary_image_raw(f: *const heic_b200_file) -> *const heic_image_desc;
    pub fn heic_b200_file_aux_image_raw(f: *const heic_b200_file, i: u32) -> *const heic_image_desc;
    pub fn heic_b200_file_parameter_set_nal(
        f: *const heic_b200_file, image: i32, nal_unit_type: u32, data: *mut *const u8, len: *mut usize,
    ) -> i32;
    pub fn heic_b200_file_tile_nal(
        f: *const heic_b200_file, image: i32, tile: u32, data: *mut *const u8, len: *mut usize,
    ) -> i32;
    pub fn heic_b200_file_info(f: *const heic_b200_file, out: *mut heic_file_info) -> i32;
    // ---- the hot path ----
    pub fn heic_b200_decode_grids(
        ctx: *mut heic_b200_ctx, imgs: *const heic_image_desc, n_imgs: u32, rgb_out: *mut u8, pitch: usize,
        image_stride: usize, apply_transforms: c_int, status: *mut heic_tile_status,
    ) -> i32;
    pub fn heic_b200_decode_grids_submit(
        ctx: *mut heic_b200_ctx, imgs: *const heic_image_desc, n_imgs: u32, rgb_out: *mut u8, pitch: usize,
        image_stride: usize, apply_transforms: c_int, status: *mut heic_tile_status, out_job: *mut *mut heic_b200_job,
    ) -> i32;
    pub fn heic_b200_job_wait(job: *mut heic_b200_job) -> i32;
    pub fn heic_b200_decode_grids_yuv(
        ctx: *mut heic_b200_ctx, imgs: *const heic_image_desc, n_imgs: u32, y_out: *mut u8, cb_out: *mut u8,
        cr_out: *mut u8, status: *mut heic_tile_status,
    ) -> i32;
    pub fn heic_b200_decode_file(
        ctx: *mut heic_b200_ctx, data: *const u8, len: usize, rgb_out: *mut u8, pitch: usize, apply_transforms: c_int,
    ) -> i32;
    // ---- resident batches ----
    pub fn heic_b200_batch_create(
        ctx: *mut heic_b200_ctx, imgs: *const heic_image_desc, n_imgs: u32, out: *mut *mut heic_b200_batch,
    ) -> i32;
    pub fn heic_b200_batch_destroy(b: *mut heic_b200_batch);
    pub fn heic_b200_batch_decode(b: *mut heic_b200_batch) -> i32;
    pub fn heic_b200_batch_run_stages(b: *mut heic_b200_batch, stage_mask: u32) -> i32;
    pub fn heic_b200_batch_sync(b: *mut heic_b200_batch) -> i32;
    pub fn heic_b200_batch_stream(b: *mut heic_b200_batch) -> *mut c_void;
    pub fn heic_b200_batch_rgb(
        b: *mut heic_b200_batch, dev_ptr: *mut *mut c_void, pitch: *mut usize, image_stride: *mut usize,
    ) -> i32;
    pub fn heic_b200_batch_download_rgb(b: *mut heic_b200_batch, rgb_out: *mut u8, pitch: usize, image_stride: usize) -> i32;
    pub fn heic_b200_batch_status(b: *mut heic_b200_batch, status: *mut heic_tile_status) -> i32;
    pub fn heic_b200_batch_download_image(b: *mut heic_b200_batch, image_index: u32, rgb_out: *mut u8, pitch: usize) -> i32;
    pub fn heic_b200_batch_tile_count(b: *const heic_b200_batch) -> u32;
    pub fn heic_b200_batch_cabac_order(b: *const heic_b200_batch, out: *mut u32, cap: usize, tiles_per_group: *mut u32) -> usize;
    pub fn heic_b200_batch_dump_tile(b: *mut heic_b200_batch, tile_index: u32, dump: *mut heic_tile_dump) -> i32;
    // ---- stand-alone stages on caller-owned buffers ----
    pub fn heic_b200_unescape(
        ctx: *mut heic_b200_ctx, nal_payload: *const u8, len: usize, slice_data_byte_offset: u32,
        substream_offset: *const u32, n_substreams: u32, rbsp_out: *mut u8, rbsp_len: *mut usize,
        slice_data_byte_offset_out: *mut u32, substream_offset_out: *mut u32,
    ) -> i32;
    pub fn heic_b200_color_stitch(
        ctx: *mut heic_b200_ctx, dev_planes: *const c_void, n_images: u32, grid_rows: u32, grid_cols: u32, tile_w: u32,
        tile_h: u32, out_w: u32, out_h: u32, full_range: u32, matrix_coeffs: u32, dev_rgb: *mut c_void, pitch: usize,
        image_stride: usize,
    ) -> i32;
}

/// One decoder = one CUDA stream on one device.  Not `Sync`: use one per thread.
pub struct B200Decoder(*mut heic_b200_ctx);

impl B200Decoder {
    pub fn new(device: i32) -> Result<Self, String> {
        let mut ctx = std::ptr::null_mut();
        let rc = unsafe { heic_b200_create(device, &mut ctx) };
        if rc < 0 { Err(last_error()) } else { Ok(Self(ctx)) }
    }

    /// Replacement for the reference's tile loop: all tiles of all images in one call; returns interleaved RGB8.
    pub fn decode_grids(&mut self, imgs: &[heic_image_desc], width: usize, height: usize) -> Result<Vec<u8>, String> {
        let mut rgb = vec![0u8; imgs.len() * width * height * 3];
        let rc = unsafe {
            heic_b200_decode_grids(self.0, imgs.as_ptr(), imgs.len() as u32, rgb.as_mut_ptr(), width * 3,
                                   width * height * 3, 0, std::ptr::null_mut())
        };
        if rc < 0 { Err(last_error()) } else { Ok(rgb) }
    }
}

/// A submitted `decode_grids` call (the asynchronous form).  The output buffer is owned by the
/// job until `wait` hands it back, so it cannot be read or dropped while the device still writes it.
pub struct B200Job {
    job: *mut heic_b200_job,
    rgb: Vec<u8>,
}

impl B200Job {
    /// Blocks until the RGB of this call is complete and returns it.
    pub fn wait(mut self) -> Result<Vec<u8>, String> {
        let rc = unsafe { heic_b200_job_wait(self.job) };
        self.job = std::ptr::null_mut();
        let rgb = std::mem::take(&mut self.rgb);
        if rc < 0 { Err(last_error()) } else { Ok(rgb) }
    }
}

impl Drop for B200Job {
    fn drop(&mut self) {
        if !self.job.is_null() {
            unsafe { heic_b200_job_wait(self.job) }; // the buffer must outlive the copies
        }
    }
}

impl B200Decoder {
    /// Queues all copies and kernels of one call and returns; submit the next call before waiting for this one to
    /// overlap its host->device copy and kernels with this call's device->host copy.
    pub fn submit_grids(&mut self, imgs: &[heic_image_desc], width: usize, height: usize) -> Result<B200Job, String> {
        let mut rgb = vec![0u8; imgs.len() * width * height * 3];
        let mut job = std::ptr::null_mut();
        let rc = unsafe {
            heic_b200_decode_grids_submit(self.0, imgs.as_ptr(), imgs.len() as u32, rgb.as_mut_ptr(), width * 3,
                                          width * height * 3, 0, std::ptr::null_mut(), &mut job)
        };
        if rc < 0 { Err(last_error()) } else { Ok(B200Job { job, rgb }) }
    }

    /// `HeicDecoder::decode(&[u8])` that returns the image (decoder.rs:12 returns `()`): (width, height, RGB8).
    pub fn decode_file(&mut self, data: &[u8]) -> Result<(u32, u32, Vec<u8>), String> {
        let mut f = std::ptr::null_mut();
        if unsafe { heic_b200_file_open(data.as_ptr(), data.len(), &mut f) } < 0 {
            return Err(last_error());
        }
        let mut info = heic_file_info::default();
        let rc = unsafe { heic_b200_file_info(f, &mut info) };
        unsafe { heic_b200_file_close(f) };
        if rc < 0 {
            return Err(last_error());
        }
        let (w, h) = (info.ispe_width as usize, info.ispe_height as usize);
        let mut rgb = vec![0u8; w * h * 3];
        let rc = unsafe { heic_b200_decode_file(self.0, data.as_ptr(), data.len(), rgb.as_mut_ptr(), w * 3, 0) };
        if rc < 0 { Err(last_error()) } else { Ok((info.ispe_width, info.ispe_height, rgb)) }
    }
}

impl Drop for B200Decoder {
    fn drop(&mut self) {
        unsafe { heic_b200_destroy(self.0) }
    }
}

fn last_error() -> String {
    unsafe { std::ffi::CStr::from_ptr(heic_b200_last_error()).to_string_lossy().into_owned() }
}
