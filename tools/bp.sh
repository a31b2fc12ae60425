cd tools
./bulk_probe 29.36 4096 8 6 16 1 148 0 0
./bulk_probe 29.36 4096 8 6 16 1 148 0 64
./bulk_probe 29.36 4096 8 6 16 1 148 0 16
./bulk_probe 29.36 4096 8 6 16 1 148 0 128
./bulk_probe 29.36 4096 8 6 16 1 148 1 64
./bulk_probe 29.36 4096 8 6 16 1 148 1 0
