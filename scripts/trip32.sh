timeout 300 python scripts/bench_layers.py --only attn_p3,attn_p4,c3x3_128_160,c1x1_128_128_320,ff0_256_1024_160,c3x3_512_40 2>&1 | tail -6
