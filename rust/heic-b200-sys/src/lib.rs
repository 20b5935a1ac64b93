//! SOURCE ONLY (never compiled here).  `extern "C"` declarations for `include/heic_b200.h`, limited to what the
//! reference's decode path needs, plus the safe wrapper that replaces the tile loop of
//! `HeicDecoder::decode` (src/heic/decoder.rs:98-119) and `SliceSegmentReader::read_data` (src/hevc/slice.rs:206).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int};

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

#[repr(C)]
pub struct heic_b200_ctx {
    _private: [u8; 0],
}

extern "C" {
    pub fn heic_b200_abi_version() -> i32;
    pub fn heic_b200_last_error() -> *const c_char;
    pub fn heic_b200_create(device: i32, out_ctx: *mut *mut heic_b200_ctx) -> i32;
    pub fn heic_b200_destroy(ctx: *mut heic_b200_ctx);
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
    pub fn heic_b200_decode_grids(
        ctx: *mut heic_b200_ctx, imgs: *const heic_image_desc, n_imgs: u32, rgb_out: *mut u8, pitch: usize,
        image_stride: usize, apply_transforms: c_int, status: *mut heic_tile_status,
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

impl Drop for B200Decoder {
    fn drop(&mut self) {
        unsafe { heic_b200_destroy(self.0) }
    }
}

fn last_error() -> String {
    unsafe { std::ffi::CStr::from_ptr(heic_b200_last_error()).to_string_lossy().into_owned() }
}
