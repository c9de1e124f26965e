#!/bin/bash
# usage: tools/sass_fn.sh <obj|so> <mangled-substring>   -> SASS of the first matching kernel (address + instruction),
#        e.g. tools/sass_fn.sh birp_cpp_b200/csrc/build/birp_kernels.o march_cube_kernelILb1ELi0ELi1ELi4E | python tools/sass_loops.py /dev/stdin
obj=$1; pat=$2
fn=$(cuobjdump -elf $obj 2>/dev/null | grep -o "_ZN[A-Za-z0-9_]*${pat}[A-Za-z0-9_]*" | sort -u | head -1)
cuobjdump -sass -fun "$fn" $obj | grep -E '^\s+/\*[0-9a-f]{4}\*/' | sed 's/\/\* 0x[0-9a-f]* \*\///; s/ *$//'
