from .._unet2d import (AttnDownBlock2D, AttnUpBlock2D, DownBlock2D, UNetMidBlock2D, UpBlock2D,  # noqa: F401
                       get_down_block, get_up_block)
