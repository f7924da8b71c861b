mkdir -p gpurun_out
L=gpurun_out/attn_prof2.log; : > $L
SKB_ATT_NQ=1 SKB_ATT_PROF=1 timeout 300 python scripts/attn_prof.py >> $L 2>&1
SKB_ATT_NQ=1 SKB_ATT_PROF=1 SKB_ATT_ONE=1 timeout 300 python scripts/attn_prof.py >> $L 2>&1
SKB_ATT_NQ=1 SKB_ATT_PROF=1 SKB_ATT_ONE=1 SKB_ATT_DBG=2 timeout 300 python scripts/attn_prof.py >> $L 2>&1
SKB_ATT_NQ=1 SKB_ATT_PROF=1 SKB_ATT_DBG=2 timeout 300 python scripts/attn_prof.py >> $L 2>&1
SKB_ATT_NQ=1 SKB_ATT_PROF=1 SKB_ATT_DBG=1 timeout 300 python scripts/attn_prof.py >> $L 2>&1
SKB_ATT_NQ=2 SKB_ATT_PROF=1 timeout 300 python scripts/attn_prof.py >> $L 2>&1
cat $L
