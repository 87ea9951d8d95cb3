#!/bin/bash
# usage: tools/sass_of.sh <kernel-name-substring>   — SASS of one kernel of liborbc_b200.so, cleaned
cuobjdump -sass "$(dirname "$0")/../openrbc_b200/liborbc_b200.so" | awk -v pat="$1" '/Function :/{p = index($0, pat) > 0} p' | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed 's/ *\/\* 0x[0-9a-f]* \*\///' | awk '{$1=$1};1'
