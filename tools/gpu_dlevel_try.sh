cd /root/repo; mkdir -p gpurun_out
bash tools/variants.sh "-DCOUNT_THREADS=512" "-DCOUNT_UNROLL_PAIRS" "-DCOUNT_UNROLL_PAIRS -DCOUNT_THREADS=256 -DCOUNT_MIN_CTAS=4"
