#!/bin/bash
# tools/sass_fn.sh <object> <mangled-kernel-substring>  -- SASS of one kernel (development aid)
cuobjdump -sass "$1" | awk -v k="$2" '/Function : /{a = index($0, k) > 0} a'
