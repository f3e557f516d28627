cd /root/repo
MPP_B200_DEBUG=1 MPP_B200_DEBUG_LIB=tools/_trace_build.so timeout 300 python tools/visit_timers.py > gpurun_out/dbg_time.log 2>&1
echo done
