mkdir -p gpurun_out
for k in text mixed; do echo "#### $k"; MIB=1024 KIND=$k bash scripts/gpu_ab_libs.sh "5:1904" lib_d256_4.so lib_d128_9.so lib_d192_6.so lib_d128_10.so lib_d256_5.so; done
