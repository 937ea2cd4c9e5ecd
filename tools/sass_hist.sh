#!/bin/bash
# usage: tools/sass_hist.sh <object> <kernel-name-substring>   -- opcode histogram of one kernel (cuobjdump -sass)
obj=$1; pat=$2
fn=$(cuobjdump -sass "$obj" 2>/dev/null | grep "Function :" | sed 's/.*Function : //' | grep -- "$pat" | head -1)
echo "$fn"
cuobjdump -sass -fun "$fn" "$obj" | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+(@!?U?P[0-9T]+\s+)?([A-Z0-9_.]+).*/\2/' | sed -E 's/^([A-Z0-9]+(\.WIDE)?(\.HI)?).*/\1/' | sort | uniq -c | sort -rn | head -${3:-16}
