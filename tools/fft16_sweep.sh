#!/bin/bash
# FFT path A/B runs per phase (run on the GPU box)
cd "$(dirname "$0")/.."
run() { echo "== $*"; env "$@" python tools/fft_phase.py 2>&1 | grep "us per"; }
run B200CA_FFT_ONLY=0
run B200CA_FFT_ONLY=3
run B200CA_FFT_ONLY=3 B200CA_FFT_FUSED=4
run B200CA_FFT_ONLY=3 B200CA_FFT_FUSED=4 B200CA_FFT_MINB=3
run B200CA_FFT_ONLY=3 B200CA_FFT_FUSED=8 B200CA_FFT_MINB=3
