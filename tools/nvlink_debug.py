"""prints what NVML exposes for the NVLink throughput counters on this box (debug aid for bench.py's NvlinkCounter)"""
import pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
for name in ("NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX", "NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX", "NVML_FI_DEV_NVLINK_THROUGHPUT_RAW_TX", "NVML_FI_DEV_NVLINK_THROUGHPUT_RAW_RX"):
    fid = getattr(pynvml, name)
    for scope in (0xFFFFFFFF, 0, 1):
        try:
            v = pynvml.nvmlDeviceGetFieldValues(h, [(fid, scope)])[0]
            print(name, hex(scope), "ret", v.nvmlReturn, "type", v.valueType, "ull", v.value.ullVal, "ul", v.value.ulVal)
        except Exception as e:  # noqa: BLE001
            print(name, hex(scope), "EXC", repr(e))
    try:
        v = pynvml.nvmlDeviceGetFieldValues(h, [fid])[0]
        print(name, "plain", "ret", v.nvmlReturn, "type", v.valueType, "ull", v.value.ullVal)
    except Exception as e:  # noqa: BLE001
        print(name, "plain EXC", repr(e))
