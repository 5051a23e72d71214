"""Kernel instantiations each BASELINE-shape workload must dispatch to (prefixes of the names b200ca_debug_variants()
reports).  tests/test_gpu_baseline_shapes.py asserts them after its parity step, bench.py records the same log in its JSON
line (`config.kernels`), so a change of the dispatch rules shows up here first."""

EXPECT = {
    # Lenia 1024^2 (BASELINE configs[1] and its C=3 form)
    "lenia1_1k_fft": ["lenia_march<single,balanced:"],
    "lenia3_1k_fft": ["lenia_march<multi,balanced:"],
    "lenia1_1k_fft_rows": ["lenia_fft_rows<1,1024,split>", "lenia_fft_firinv<8,single,1,512>"],
    "lenia3_1k_fft_rows": ["lenia_fft_rows<2,1024,split>", "lenia_fft_firinv<4,multi,2,256>"],
    "lenia1_1k_direct": ["lenia_conv<"],
    "lenia3_1k_direct": ["lenia_conv<"],
    # MaCE-XChan (BASELINE configs[2])
    "xchan_1k": ["lenia_fft_rows<", "lenia_fft_firinv<", "mace_fast<3,"],
    "xchan_4k": ["lenia_march<multi,balanced:", "mace_fast<3,16>"],
    "xchan_4k_hybrid": ["lenia_fft_rows<2,1024,split>", "lenia_fft_firinv<8,multi,2,256>", "mace_fast<3,16>"],
    # overlap-save segments / the 16384^2 world of BASELINE configs[4]
    "lenia_seg_4k": ["lenia_march<single,balanced:"],
    "lenia_16k": ["lenia_march<single,balanced:"],
    "gol_64k": ["gol_fast<1,0,16>"],
    "nca_256_impl1": ["nca_update_tc<grid=148>"],
    "nca_256_impl0": ["nca_update_simt"],
}
