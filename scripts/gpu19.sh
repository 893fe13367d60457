cd scripts/tma_probe
for args in "64 32 8 40 11 -4 -1 1 0 2" "64 32 8 40 11 28 25 -1 0 2" "64 32 8 40 11 60 -1 7 0 2" "28 24 20 40 11 -4 -1 3 0 2" "28 24 20 40 11 28 15 19 0 2"; do echo "== $args"; ./probe $args 2>&1 | tail -2; done
