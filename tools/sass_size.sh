#!/bin/bash
# usage: tools/sass_size.sh <object-or-so>   -> .text size (bytes) of every kernel in it
d=$(mktemp -d); cd $d && cuobjdump -xelf all "$1" >/dev/null && size -A *.cubin | grep "^\.text" | awk '{print $2, $1}' | sed 's/_ZN[0-9]*_GLOBAL__N__[0-9a-f_]*cu_[0-9a-f]*//'
