python bench.py --train-only --train-dropout 0.0 > gpurun_out/train_d0.json 2> gpurun_out/train_d0.err
python bench.py --train-only --train-dropout 0.2 > gpurun_out/train_d2.json 2> gpurun_out/train_d2.err
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_train.csv python bench.py --train-only --train-dropout 0.2 --train-steps 1 --profile-range train > gpurun_out/launches_train.out 2>&1
