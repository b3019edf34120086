# A/B builds of libskani_b200.so with different compile-time knobs -> skder_b200/_lib/ab_<name>.so (git-ignored, travels to the GPU box)
# usage: tools/build_variants.sh name1 "-DSKB_X=1 ..." name2 "..." ...
set -e
cd "$(dirname "$0")/.."
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared $flags \
    -Xptxas=-v -o skder_b200/_lib/ab_$name.so skder_b200/csrc/skb_api.cu skder_b200/csrc/fasta_pack.cpp -lz 2>&1 | grep -A2 "Function properties for _ZN3skb13anchor_kernelILb1" | tail -2 | sed "s/^/[$name] /" &
done
wait
