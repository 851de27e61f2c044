# One small search through every sharded-pass variant with virtual ranks (RIES_FAKE_SHARD) on one GPU, compared with the golden text.
export HOME=/nonexistent
run() { # name, env..., then args
  name=$1; shift
  out=$(env "$@" timeout 60 ./ries_b200/bin/ries 2.5063 -l3 -F3 -x --no-slow-messages 2>&1); rc=$?
  ref=$(sed '$d' tests/golden/t25063_l3_F3_x.txt)
  got=$(echo "$out" | sed -E 's/Time: [0-9.]+/Time: X/')
  if [ "$got" == "$ref" ]; then echo "$name rc=$rc SAME"; else echo "$name rc=$rc DIFF"; echo "$got" | diff - <(echo "$ref") | head -8; fi
}
run plain A=1
run fake4_norange RIES_FAKE_SHARD=4 RIES_SHARD_MIN=0 RIES_NO_RANGE_SPLIT=1
run fake3_range16 RIES_FAKE_SHARD=3 RIES_SHARD_MIN=0 RIES_RANGE_MIN=16
run fake2_range RIES_FAKE_SHARD=2 RIES_SHARD_MIN=2000 RIES_RANGE_MIN=1000
run fake5_events RIES_FAKE_SHARD=5 RIES_SHARD_MIN=500 RIES_RANGE_MIN=1000 RIES_EVENTS_CAP=3
run fake8_range RIES_FAKE_SHARD=8 RIES_SHARD_MIN=0 RIES_RANGE_MIN=64
