#!/bin/bash
# SASS opcode histogram of the in-tree libsvit.so (proof of tcgen05 / TMA / TMEM use).  Usage: bash scripts/sass_histogram.sh > profiles/r2_sass_histogram.md
SO=${1:-shapley_vit_b200/csrc/libsvit.so}
echo "# SASS opcode histogram of \`$SO\` (cuobjdump -sass, sm_100a)"
echo
echo "Build: \`python -m shapley_vit_b200.build\` (nvcc $(nvcc --version | grep -o 'release [0-9.]*'), -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo)."
echo
echo "## Blackwell-specific instructions (tensor cores, tensor memory, TMA, clusters)"
echo
echo '| mnemonic | count | meaning |'
echo '|---|---:|---|'
cuobjdump -sass "$SO" > /tmp/svit_sass.txt
count() { grep -c -E "^\s+/\*[0-9a-f]+\*/\s+(@!?U?P[0-9T]+ )?$1" /tmp/svit_sass.txt; }
row() { echo "| \`$1\` | $(count "$2") | $3 |"; }
row "UTCHMMA (all)" "UTCHMMA" "tcgen05.mma kind::f16 / tf32"
row "UTCHMMA.2CTA" "UTCHMMA\.2CTA" "tcgen05.mma cta_group::2 (CTA-pair MMA)"
row "UTCQMMA (all)" "UTCQMMA" "tcgen05.mma kind::f8f6f4 (e4m3 compensation passes of f16c8)"
row "UTCQMMA.2CTA" "UTCQMMA\.2CTA" "the same, CTA pair"
row "UTCBAR" "UTCBAR" "tcgen05.commit -> mbarrier"
row "LDTM" "LDTM" "tcgen05.ld (TMEM -> registers)"
row "STTM" "STTM" "tcgen05.st (registers -> TMEM)"
row "UTCATOMSWS / alloc" "UTCATOMSWS" "tcgen05.alloc / dealloc"
row "UTMALDG" "UTMALDG" "cp.async.bulk.tensor load (TMA)"
row "UTMASTG" "UTMASTG" "cp.async.bulk.tensor store (TMA)"
row "UTMAREDG" "UTMAREDG" "cp.reduce.async.bulk.tensor (TMA reduce-add into L2)"
row "UBLKCP" "UBLKCP" "cp.async.bulk (1-D bulk copy, K1)"
row "SYNCS" "SYNCS" "mbarrier operations"
row "UCGABAR" "UCGABAR" "barrier.cluster"
row "FFMA2" "FFMA2" "packed fp32x2 FMA"
row "HMMA" "HMMA" "warp-level mma.sync (T <= 128 attention fallback only)"
echo
echo "## Top 40 opcodes (base mnemonic)"
echo
echo '| opcode | count |'
echo '|---|---:|'
grep -E "^\s+/\*[0-9a-f]+\*/" /tmp/svit_sass.txt | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//; s/^@!?U?P[0-9T]+ //' | awk '{print $1}' | sed -E 's/\..*//; s/;//' | sort | uniq -c | sort -rn | head -40 | awk '{print "| `"$2"` | "$1" |"}'
echo
echo "## Kernels in the library"
echo
cuobjdump -sass "$SO" | grep -E "^\s+Function :" | sed -E 's/^\s+Function : //' | c++filt | sed -E 's/svit::\(anonymous namespace\):://; s/\(.*//' | sort | uniq -c | awk '{c=$1; $1=""; print "- `" substr($0,2) "` x" c}'
