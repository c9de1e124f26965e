#!/bin/bash
# SASS of the four default march kernels -> profiles/sass_march_*.txt (diffable round over round) + their loop lengths.
# No GPU needed.  usage: tools/sass_defaults.sh  (after the build)
cd "$(dirname "$0")/.."
obj=birp_cpp_b200/csrc/build/birp_kernels.o
while read name pat; do
  tools/sass_fn.sh $obj "$pat" > profiles/sass_march_$name.txt
  echo "$name: $(wc -l < profiles/sass_march_$name.txt) instructions; loops of 20-300 instructions (the per-sample loops):"
  python tools/sass_loops.py profiles/sass_march_$name.txt | awk '{n=$3+0; if (n >= 20 && n <= 300) print "   " $0}'
done <<LIST
cube_exact_lean march_cube_kernelILb1ELi0ELi1ELi5ELi2E
cube_exact_marked march_cube_kernelILb1ELi0ELi3ELi5ELi1E
cube_fast march_cube_fast_kernelILb0ELi2ELb1E
cmem_default march_kernelINS0_9CmemFieldILi3E
LIST
