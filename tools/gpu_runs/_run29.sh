mkdir -p gpurun_out
timeout 300 python tools/two_in_flight.py > gpurun_out/r2_two_in_flight.json 2> gpurun_out/r2_two_in_flight.err
echo done
