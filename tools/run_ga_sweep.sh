for v in "$@"; do echo "== $v"
  if [[ $v =~ s[0-9]+c ]]; then FINENVS_B200_LIB=$PWD/finenvs_b200/libfe_ga_$v.so timeout 120 python tools/gather_clocks.py 2>&1 | tail -8
  else FINENVS_B200_LIB=$PWD/finenvs_b200/libfe_ga_$v.so timeout 120 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-also 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['roofline']['kernel'], 'dev ms', round(d['ms_per_step'],4), 'e2e ms', round(d['e2e']['ms_per_step'],4))"; fi
done
