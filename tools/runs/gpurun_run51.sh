O=gpurun_out
python tools/time_poisson.py > $O/r2_s51_poisson.txt 2> $O/r2_s51_poisson.err; cat $O/r2_s51_poisson.txt; tail -5 $O/r2_s51_poisson.err
