#!/bin/bash
# Build the command-line driver faqcs_b200/host/faqcs_b200 against the in-tree C-ABI library.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
/usr/bin/g++ -O3 -std=c++17 -Wall -Wno-unused-function "$HERE/faqcs_cli.cpp" -o "$HERE/faqcs_b200" \
    -L"$HERE/.." -lfaqcs_b200 -lz -lpthread -Wl,-rpath,'$ORIGIN/..'
echo "built $HERE/faqcs_b200"
