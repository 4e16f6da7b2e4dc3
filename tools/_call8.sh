T="python tools/time_config.py 30 10000 365 1000 5"
V=abm_amro_b200/lib_variants
{
echo -n "cur: "; AMRO_B200_LIB=$V/cur/libamro_b200.so $T 2>&1 | tail -1
echo -n "cur lean: "; AMRO_FAST_LEAN=1 AMRO_B200_LIB=$V/cur/libamro_b200.so $T 2>&1 | tail -1
echo -n "lr64 lean: "; AMRO_FAST_LEAN=1 AMRO_B200_LIB=$V/lr64/libamro_b200.so $T 2>&1 | tail -1
echo -n "cur mp: "; AMRO_B200_LIB=$V/cur/libamro_b200.so python tools/time_config.py 20 5000 180 32768 3 2>&1 | tail -1
echo -n "lr64 mp: "; AMRO_B200_LIB=$V/lr64/libamro_b200.so python tools/time_config.py 20 5000 180 32768 3 2>&1 | tail -1
} > gpurun_out/sweep8.txt 2>&1
cat gpurun_out/sweep8.txt
