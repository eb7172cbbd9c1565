#!/bin/bash
cd tools/ozaki
for args in "300 500 37 0.7 12 1.0 1 2" "1800 1185 180 0.715 12 8 2 1" "1800 4740 180 0.715 12 50 2 1" "1800 592 180 0.715 12 50 3 1" "1800 592 180 0.715 12 50 3 2"; do
  echo "== $args"
  timeout 100 ./k3_i8_test $args 2>&1 | tail -5
  echo "rc=$?"
done
