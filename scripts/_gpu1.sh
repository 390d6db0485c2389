cd /root/repo
timeout 300 scripts/_bin/fft_vs_cufft 20 > gpurun_out/r02z_fft_vs_cufft.jsonl 2> gpurun_out/r02z_fft_vs_cufft.err; cat gpurun_out/r02z_fft_vs_cufft.jsonl; tail -2 gpurun_out/r02z_fft_vs_cufft.err
