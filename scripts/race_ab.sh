for v in "$@"; do
  if [ $v = new ]; then unset ASW_B200_LIB; else export ASW_B200_LIB=$PWD/stereo_matchin_b200/libasw_b200_$v.so; fi
  echo "== $v"; timeout 250 python scripts/race_probe3.py ${RUNS:-40}
done
