mkdir -p gpurun_out
python tools/measure_configs.py --iters 3 > gpurun_out/r2c_configs_field.jsonl 2> gpurun_out/r2c_err.log
CBFRRT_GRID_MIN=1 python tools/measure_configs.py --iters 3 --only cfg0,cfg1,cfg3 > gpurun_out/r2c_configs_field_gridmin1.jsonl 2>> gpurun_out/r2c_err.log
CBFRRT_TC_GROUPS_GRID=4 python tools/measure_configs.py --iters 3 --only cfg4 > gpurun_out/r2c_configs_field_gg4.jsonl 2>> gpurun_out/r2c_err.log
tail -3 gpurun_out/r2c_err.log
