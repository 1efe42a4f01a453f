#!/bin/bash
# SASS evidence of the hot kernels (per kernel: instruction count + the memory / vote / MUFU / FP64 instruction mix), so the
# presence or absence of 128-bit loads, evict-first streaming loads, TMA (UBLKCP / UTMALDG) and LDGSTS is reviewable.
LIB=${1:-instagraal_b200/libinstagraal_b200.so}
OUT=${2:-profiles/r2_sass_summary.txt}
cuobjdump -sass "$LIB" > /tmp/ig_all.sass
{
echo "SASS summary of $LIB (sm_100a), $(date -u +%F) -- cuobjdump -sass; counts per kernel"
echo "columns: kernel | instructions | LDG.E.128 | LDG.E.64 | LDG (other) | LDG .EF (evict-first) | STG | LDS | STS | VOTE/MATCH | SHFL | MUFU | F64 (DFMA/DADD/DMUL) | ATOM/RED | UBLKCP/UTMALDG/LDGSTS"
awk '
/Function : /{ if (name != "") flush(); name=$3; n=0; l128=l64=lo=ef=stg=lds=sts=vote=shfl=mufu=f64=atom=tma=0; next }
/^[ \t]+\/\*[0-9a-f]+\*\// {
  n++;
  if ($0 ~ /LDG\.E(\.[A-Z0-9_]+)*\.128/) l128++; else if ($0 ~ /LDG\.E(\.[A-Z0-9_]+)*\.64/) l64++; else if ($0 ~ / LDG/) lo++;
  if ($0 ~ /LDG\.E\.EF/) ef++;
  if ($0 ~ / STG/) stg++; if ($0 ~ / LDS/) lds++; if ($0 ~ / STS/) sts++;
  if ($0 ~ / VOTE| MATCH/) vote++; if ($0 ~ / SHFL/) shfl++; if ($0 ~ / MUFU/) mufu++;
  if ($0 ~ / DFMA| DADD| DMUL/) f64++; if ($0 ~ / ATOM| RED\.| ATOMG| ATOMS/) atom++;
  if ($0 ~ /UBLKCP|UTMALDG|LDGSTS/) tma++;
}
function flush() { printf "%-60s %6d %5d %5d %5d %5d %5d %5d %5d %5d %5d %5d %5d %5d %5d\n", substr(name,1,60), n, l128, l64, lo, ef, stg, lds, sts, vote, shfl, mufu, f64, atom, tma }
END { flush() }' /tmp/ig_all.sass | sort
echo
echo "note: no kernel uses TMA / cp.async: the streamed arrays (8-byte contact records) are consumed once, straight from"
echo "registers, by warps that each own a contiguous 512-byte piece per load instruction (LDG.E.EF.128 x 32 lanes); staging"
echo "them through shared memory would add a round trip without reuse.  See DESIGN.md section 4."
} > "$OUT"
wc -l "$OUT"
