//! src/ffi.rs — binding of include/wafer_b200.h.  UNTESTED here (no Rust toolchain in the build image).
//!
//! `Array3<R64>` created by `zeros` / `from_elem` / `from_shape_fn` is a standard-layout C-order array and `R64`
//! is a transparent wrapper around `f64`, so `as_ptr() as *const f64` is the padded buffer the library expects.
#![allow(dead_code)]
use errors::*;
use ndarray::Array3;
use noisy_float::prelude::*;
use std::ffi::CStr;
use std::os::raw::{c_char, c_int, c_void};
use std::ptr;

#[repr(C)]
pub struct WaferParams {
    pub nx: u64,
    pub ny: u64,
    pub nz: u64,
    pub ext: u32,
    pub dn: f64,
    pub dt: f64,
    pub mass: f64,
    pub device: i32,
    pub rank: u32,
    pub world: u32,
    pub nccl_id: *const u8,
    pub max_lower: u32,
    pub flags: u32,
}

#[repr(C)]
#[derive(Default, Debug, Clone, Copy)]
pub struct WaferObservables {
    pub energy: f64,
    pub norm2: f64,
    pub v_infinity: f64,
    pub r2: f64,
}

pub enum WaferCtx {}

extern "C" {
    fn wafer_create(p: *const WaferParams, out: *mut *mut WaferCtx) -> c_int;
    fn wafer_destroy(ctx: *mut WaferCtx) -> c_int;
    fn wafer_last_error(ctx: *const WaferCtx) -> *const c_char;
    fn wafer_set_potential(ctx: *mut WaferCtx, v_padded: *const f64) -> c_int;
    fn wafer_set_pot_sub_scalar(ctx: *mut WaferCtx, c: f64) -> c_int;
    fn wafer_set_pot_sub_array(ctx: *mut WaferCtx, work: *const f64) -> c_int;
    fn wafer_set_phi(ctx: *mut WaferCtx, phi_padded: *const f64) -> c_int;
    fn wafer_get_phi(ctx: *mut WaferCtx, phi_padded: *mut f64) -> c_int;
    fn wafer_push_lower(ctx: *mut WaferCtx, q_padded: *const f64) -> c_int;
    fn wafer_push_lower_from_phi(ctx: *mut WaferCtx) -> c_int;
    fn wafer_phi_from_lower(ctx: *mut WaferCtx, idx: u32) -> c_int;
    fn wafer_phi_seed_from_lower(ctx: *mut WaferCtx, idx: u32) -> c_int;
    fn wafer_check(ctx: *mut WaferCtx, wnum: u8, out: *mut WaferObservables) -> c_int;
    fn wafer_normalise(ctx: *mut WaferCtx, norm2: f64) -> c_int;
    fn wafer_evolve(ctx: *mut WaferCtx, wnum: u8, steps: u64) -> c_int;
    fn wafer_synchronize(ctx: *mut WaferCtx) -> c_int;
    fn wafer_host_register(ptr: *mut c_void, bytes: usize) -> c_int;
    fn wafer_host_unregister(ptr: *mut c_void) -> c_int;
}

/// Owns one `wafer_ctx` (one GPU).  Used from the single thread that runs `grid::solve`.
pub struct Gpu {
    ctx: *mut WaferCtx,
}

impl Gpu {
    fn check(&self, rc: c_int) -> Result<()> {
        if rc == 0 {
            return Ok(());
        }
        let msg = unsafe { CStr::from_ptr(wafer_last_error(self.ctx)).to_string_lossy().into_owned() };
        if rc == 7 {
            Err(ErrorKind::MaxStep.into())
        } else {
            Err(format!("GPU hot path failed (status {}): {}", rc, msg).into())
        }
    }

    pub fn new(size: (usize, usize, usize), ext: usize, dn: R64, dt: R64, mass: R64, wavemax: u8) -> Result<Gpu> {
        let p = WaferParams {
            nx: size.0 as u64,
            ny: size.1 as u64,
            nz: size.2 as u64,
            ext: ext as u32,
            dn: dn.raw(),
            dt: dt.raw(),
            mass: mass.raw(),
            device: -1,
            rank: 0,
            world: 1,
            nccl_id: ptr::null(),
            max_lower: wavemax as u32,
            flags: 0,
        };
        let mut ctx: *mut WaferCtx = ptr::null_mut();
        let rc = unsafe { wafer_create(&p, &mut ctx) };
        if rc != 0 {
            let msg = unsafe { CStr::from_ptr(wafer_last_error(ptr::null())).to_string_lossy().into_owned() };
            return Err(format!("wafer_create failed (status {}): {}", rc, msg).into());
        }
        Ok(Gpu { ctx })
    }

    pub fn set_potential(&self, v: &Array3<R64>) -> Result<()> {
        self.check(unsafe { wafer_set_potential(self.ctx, v.as_ptr() as *const f64) })
    }
    pub fn set_pot_sub(&self, pot_sub: &(Option<Array3<R64>>, Option<R64>)) -> Result<()> {
        match *pot_sub {
            (Some(ref arr), None) => self.check(unsafe { wafer_set_pot_sub_array(self.ctx, arr.as_ptr() as *const f64) }),
            (None, Some(c)) => self.check(unsafe { wafer_set_pot_sub_scalar(self.ctx, c.raw()) }),
            _ => self.check(unsafe { wafer_set_pot_sub_scalar(self.ctx, 0.0) }),
        }
    }
    pub fn set_phi(&self, phi: &Array3<R64>) -> Result<()> {
        self.check(unsafe { wafer_set_phi(self.ctx, phi.as_ptr() as *const f64) })
    }
    pub fn get_phi(&self, phi: &mut Array3<R64>) -> Result<()> {
        self.check(unsafe { wafer_get_phi(self.ctx, phi.as_mut_ptr() as *mut f64) })
    }
    pub fn push_lower(&self, q: &Array3<R64>) -> Result<()> {
        self.check(unsafe { wafer_push_lower(self.ctx, q.as_ptr() as *const f64) })
    }
    pub fn push_lower_from_phi(&self) -> Result<()> {
        self.check(unsafe { wafer_push_lower_from_phi(self.ctx) })
    }
    pub fn phi_seed_from_lower(&self, idx: u32) -> Result<()> {
        self.check(unsafe { wafer_phi_seed_from_lower(self.ctx, idx) })
    }
    pub fn check_state(&self, wnum: u8) -> Result<WaferObservables> {
        let mut o = WaferObservables::default();
        self.check(unsafe { wafer_check(self.ctx, wnum, &mut o) })?;
        Ok(o)
    }
    pub fn normalise(&self, norm2: f64) -> Result<()> {
        self.check(unsafe { wafer_normalise(self.ctx, norm2) })
    }
    pub fn evolve(&self, wnum: u8, steps: u64) -> Result<()> {
        self.check(unsafe { wafer_evolve(self.ctx, wnum, steps) })
    }
    /// Page-lock the buffer behind `arr` so that set_phi / get_phi run at PCIe speed (54 GB/s measured instead of the
    /// driver's 14 GB/s pageable path).  Call once after allocating the array; `unpin` before it is dropped.
    pub fn pin(&self, arr: &mut Array3<R64>) -> Result<()> {
        let bytes = arr.len() * ::std::mem::size_of::<f64>();
        self.check(unsafe { wafer_host_register(arr.as_mut_ptr() as *mut c_void, bytes) })
    }
    pub fn unpin(&self, arr: &mut Array3<R64>) {
        unsafe { wafer_host_unregister(arr.as_mut_ptr() as *mut c_void) };
    }
}

impl Drop for Gpu {
    fn drop(&mut self) {
        unsafe {
            wafer_synchronize(self.ctx);
            wafer_destroy(self.ctx);
        }
    }
}
