//! GPU blocks with the block signature of LibRedio's `kpn` crate
//! (`pub fn name(u: Receiver<In>, v: Sender<Out>, params…)`, src/kpn/src/kpn.rs:127-131).
//! Each spawns on its own thread exactly like the blocks of src/ratpak.rs:60-185; a closed port or a
//! non-zero status of the GPU library panics the block, which drops its ports and tears the graph down --
//! the reference's own convention (`.unwrap()` everywhere, `panic!(src_strerror(..))` samplerate.rs:77-83).
//!
//! SOURCE ONLY in this repository (no Rust toolchain in the build image); the compiled and tested host
//! layer is the C++ mirror kpn/gpu_blocks.hpp.  The host-buffer entry points of the C ABI are used so this
//! crate needs no CUDA bindings of its own.
use libredio_cuda_sys as sys;
use num_complex::Complex;
use std::ffi::CStr;
use std::ptr;
use std::sync::mpsc::{Receiver, Sender};

fn check(rc: i32, what: &str) {
    if rc != sys::LRC_OK {
        let detail = unsafe { CStr::from_ptr(sys::lrc_last_error()) }.to_string_lossy().into_owned();
        panic!("{}: {}", what, detail);
    }
}

/// One context per GPU; share it between the blocks of a graph.
pub struct Gpu(pub *mut sys::lrc_ctx);
unsafe impl Send for Gpu {}
unsafe impl Sync for Gpu {}
impl Gpu {
    pub fn new(device: i32) -> Gpu {
        let mut ctx = ptr::null_mut();
        check(unsafe { sys::lrc_ctx_create(device, &mut ctx) }, "lrc_ctx_create");
        Gpu(ctx)
    }
}
impl Drop for Gpu {
    fn drop(&mut self) { unsafe { sys::lrc_ctx_destroy(self.0); } }
}

/// Drop-in for `kissfft::fft(pin, cout, block_size, inv)` (src/kissfft/src/kissfft.rs:18-31).
/// Every frame already queued on `pin` goes to the device in one launch.
pub fn fft(gpu: &Gpu, pin: Receiver<Vec<Complex<f32>>>, cout: Sender<Vec<Complex<f32>>>, block_size: u32, inv: u32) {
    let mut plan = ptr::null_mut();
    check(unsafe { sys::lrc_fft_create(gpu.0, block_size as i32, inv as i32, &mut plan) }, "lrc_fft_create");
    let n = block_size as usize;
    loop {
        let mut frames = vec![pin.recv().unwrap()];
        while let Ok(f) = pin.try_recv() { frames.push(f); if frames.len() >= 4096 { break; } }
        let mut flat: Vec<Complex<f32>> = Vec::with_capacity(frames.len() * n);
        for f in &frames {
            assert!(f.len() == n);                       // kissfft.rs:24
            flat.extend_from_slice(f);
        }
        let mut out = vec![Complex::new(0f32, 0f32); flat.len()];
        check(unsafe { sys::lrc_fft_run_host(plan, flat.as_ptr() as *const f32, out.as_mut_ptr() as *mut f32, flat.len()) },
              "lrc_fft_run_host");
        for k in 0..frames.len() { cout.send(out[k * n..(k + 1) * n].to_vec()).unwrap(); }
    }
}

/// Drop-in for `samplerate::resample(din, dout, ratio)` (src/samplerate/src/samplerate.rs:59-87).
pub fn resample(gpu: &Gpu, din: Receiver<Vec<f32>>, dout: Sender<Vec<f32>>, ratio: f64) {
    let mut rs = ptr::null_mut();
    check(unsafe { sys::lrc_resampler_create(gpu.0, ratio, 1, 1 << 22, &mut rs) }, "lrc_resampler_create");
    loop {
        let vin = din.recv().unwrap();
        let lout = ((ratio * vin.len() as f64) + 1f64) as usize + 1;      // samplerate.rs:64
        let mut vout = vec![0f32; lout];
        let mut n_out = 0usize;
        check(unsafe { sys::lrc_resampler_process_host(rs, vin.as_ptr(), vin.len(), vout.as_mut_ptr(), lout, &mut n_out) },
              "lrc_resampler_process_host");
        vout.truncate(n_out);                                             // set_len(output_frames_gen) :84
        dout.send(vout).unwrap();
    }
}

/// The headline chain as one block: cf32 chunks holding whole rows of `k_avg` frames (+ the ntaps-decim tail)
/// in, one `Vec<f32>` of `nfft` averaged |X|^2 values per row out.
pub fn chain_psd(gpu: &Gpu, u: Receiver<Vec<Complex<f32>>>, v: Sender<Vec<f32>>, taps: &[f32], decim: usize, nfft: usize, k_avg: usize) {
    let mut ch = ptr::null_mut();
    check(unsafe { sys::lrc_chain_create(gpu.0, taps.as_ptr(), taps.len() as i32, decim as i32, nfft as i32, sys::LRC_WINDOW_HANN, &mut ch) },
          "lrc_chain_create");
    loop {
        let x = u.recv().unwrap();
        let rows = unsafe { sys::lrc_chain_frames(ch, x.len()) } / k_avg;
        let mut out = vec![0f32; rows * nfft];
        let mut nr = 0usize;
        check(unsafe { sys::lrc_chain_run_host(ch, x.as_ptr() as *const f32, x.len(), k_avg, out.as_mut_ptr(), &mut nr) },
              "lrc_chain_run_host");
        for r in 0..nr { v.send(out[r * nfft..(r + 1) * nfft].to_vec()).unwrap(); }
    }
}

/// `kpn::eat` (src/kpn/src/kpn.rs:116-124) served by the library's host helper.
pub fn eat(x: &[usize], is: &[usize]) -> Vec<usize> {
    let bits: Vec<u8> = x.iter().map(|&b| b as u8).collect();
    let mut out = vec![0usize; is.len()];
    check(unsafe { sys::lrc_eat(bits.as_ptr(), bits.len(), is.as_ptr(), is.len(), out.as_mut_ptr()) }, "lrc_eat");
    out
}
