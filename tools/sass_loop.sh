#!/bin/bash
# Dump the SASS of the interior profile kernel and print the size of its march loop (WARPSYNC.ALL #2 .. #3).
SO=${1:-smarc_navigation_b200/libauvpf.so}
cuobjdump -sass $SO | awk '/Function : .*k_update_mbes_profileILi128ELi7ELb0/{f=1} /Function : /{if(f&&!/k_update_mbes_profileILi128ELi7ELb0/)f=0} f{print}' | grep -v "^\s*/\* 0x" | sed 's#/\* 0x[0-9a-f]* \*/##' > /tmp/prof_sass.txt
A=$(grep -n "WARPSYNC.ALL" /tmp/prof_sass.txt | sed -n 2p | cut -d: -f1); B=$(grep -n "WARPSYNC.ALL" /tmp/prof_sass.txt | sed -n 3p | cut -d: -f1)
echo "loop lines $A..$B = $((B-A)) instr (static, both beam loops and side switch included)"
