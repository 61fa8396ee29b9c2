for c in gram_64_f16 gram_128_bf16 gram_256_f16 wgrad_bb_64_128 wgrad_hh_64_128 wgrad_hb_64_128 wgrad_bb_256_256 wgrad_hb_128_64; do
  timeout 120 python tools/dbg_wgrad.py $c 2>&1 | tail -2
done
