for i in 1 2; do
python tools/micro_conv.py 16 256 256 64 128 3 2 1 0 dgrad
FPG_S2CLS_ROWSTORE=1 python tools/micro_conv.py 16 256 256 64 128 3 2 1 0 dgrad
python tools/micro_conv.py 32 256 256 12 64 4 2 1 0 dgrad
FPG_S2CLS_ROWSTORE=1 python tools/micro_conv.py 32 256 256 12 64 4 2 1 0 dgrad
done
