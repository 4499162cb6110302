"""Single-GPU training.  Same command line as the reference's train_single.py (flags at :22-56)."""
from train import main

if __name__ == '__main__':
    main(multi_gpu=False)
