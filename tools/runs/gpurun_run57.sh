O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_mlpoisson.py -x -q > $O/r2_s57_pytest_ml.log 2>&1; tail -15 $O/r2_s57_pytest_ml.log
for n in 64 128 256 512; do
  timeout 120 python tools/time_mlpoisson.py $n f32 8 >> $O/r2_s57_ml.jsonl 2>> $O/r2_s57_ml.err
  IFADV_ML_GRAPH=0 timeout 120 python tools/time_mlpoisson.py $n f32 8 >> $O/r2_s57_ml.jsonl 2>> $O/r2_s57_ml.err
done
timeout 120 python tools/time_mlpoisson.py 256 f64 8 >> $O/r2_s57_ml.jsonl 2>> $O/r2_s57_ml.err
cat $O/r2_s57_ml.jsonl | cut -c 1-1200; tail -5 $O/r2_s57_ml.err
