#!/bin/bash
# tools/build_variant.sh <name> <nvcc -D flags...>: an experiment build of the library, ultrare_b200/csrc/_obj/var_<name>.so
# (select it with URE_LIB=ultrare_b200/csrc/_obj/var_<name>.so); FILES="a b": the sources to recompile with the flags
# (default mf_train_owner), the other objects are taken from the regular build
set -e
cd "$(dirname "$0")/.."
name=$1; shift
D=ultrare_b200/csrc/_obj/var_$name; mkdir -p $D
for f in ultrare_b200/csrc/*.cu; do
  b=$(basename $f .cu)
  if [[ " ${FILES:-mf_train_owner} " == *" $b "* ]] || [ ! -f ultrare_b200/csrc/_obj/$b.o ]; then
    /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c $f -o $D/$b.o &
  else
    cp ultrare_b200/csrc/_obj/$b.o $D/$b.o
  fi
done
wait
/usr/local/cuda/bin/nvcc -shared -o ultrare_b200/csrc/_obj/var_$name.so $D/*.o -lcudart
echo ultrare_b200/csrc/_obj/var_$name.so
