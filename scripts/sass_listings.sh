#!/bin/bash
# Full SASS listings of the in-tree extension, one file per kernel family (cuobjdump -sass, the
# encoding column stripped):   bash scripts/sass_listings.sh   ->  profiles/sass/listing_*.sass
# plus the per-kernel count of Blackwell-specific mnemonics (scripts/sass_evidence.sh).
set -e
OUT=profiles/sass
mkdir -p $OUT
dump() {  # $1 = object file, $2 = output name, $3 = optional function regex
  cuobjdump -sass build/tdp_b200/$1 2>/dev/null | grep -v '^\s*/\* 0x' | sed -E 's|\s*/\* 0x[0-9a-f]+ \*/\s*$||' \
    | awk -v re="${3:-.}" '/Function :/ { keep = ($0 ~ re) } keep' > $OUT/$2
  echo "$2: $(grep -c 'Function :' $OUT/$2) kernels, $(wc -l < $OUT/$2) lines"
}
dump attn_attn_fwd_sm100.cu.o      listing_attention_fwd.sass
dump attn_attn_bwd_dq_sm100.cu.o   listing_attention_bwd_dq.sass
dump attn_attn_bwd_sm100.cu.o      listing_attention_bwd_dkv.sass
dump coll_collectives.cu.o         listing_collectives_nvls.sass 'Lb1E|barrier_only|a2a_'
dump fused_optim.cu.o              listing_optim_adamw_ema_norm.sass 'adamw_kernelILb1ELb1ELb1E|ema_multi|sumsq_multi|scale_multi'
dump fused_norm_loss.cu.o          listing_layernorm_ce_colsum.sass
dump fused_layout.cu.o             listing_layout_rows_copy.sass
dump gemm_gemm.cu.o                listing_gemm_2cta_ring_epilogue.sass '2cta_kernelILi256ELi6E'
dump gemm_gemm.cu.o                listing_gemm_2cta_256.sass '2cta_kernelILi256ELi2E'
dump gemm_gemm.cu.o                listing_gemm_1cta_256.sass 'gemm_bf16_sm100_kernelILi256E'
bash scripts/sass_evidence.sh > $OUT/blackwell_mnemonics_by_kernel.txt
wc -l $OUT/blackwell_mnemonics_by_kernel.txt
