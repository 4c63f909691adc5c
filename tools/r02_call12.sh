for dbg in 0 1 2 3; do
echo "== SEUNET_CONV_DBG=$dbg"
SEUNET_CONV_DBG=$dbg CONV_BENCH_ITERS=1 CONV_BENCH_WARMUP=1 SEUNET_LIB_PATH=tools/libseunet_prof.so python tools/conv_layer_bench.py 7 128 dc5,ec3,ec2 2>&1 | awk '/conv prof/{l=$0} !/conv prof/{print l}' | sed 's/producer.*issuer/issuer/'
done
