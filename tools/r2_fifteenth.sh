# Round 2, fifteenth GPU call (1 GPU): Bluestein factorisation / pipelined-flavour knobs at 1,000,003; sanitizer over the round-2 kernels.
python tools/ab_headline.py 32 1000003
SFC_BLUE_L1=1024 python tools/ab_headline.py 32 1000003
SFC_PIPE=2 python tools/ab_headline.py 32 1000003
SFC_PIPE=1 python tools/ab_headline.py 32 1000003
SFC_BLUE_L1=1024 SFC_PIPE=2 python tools/ab_headline.py 32 1000003
bash tools/sanitize3.sh
