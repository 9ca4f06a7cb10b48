"""Host-side mirror of `diffusers.image_processor.VaeImageProcessor` as the reference uses it
(`train.py:33,839` / `app.py:13,67`: `VaeImageProcessor().preprocess(pil_image, height, width)`), SURVEY 8(f).2.

Pure host code (PIL + numpy): decode -> resize (Lanczos) to a multiple of the VAE scale factor -> [0, 1] -> [-1, 1],
NCHW float32, exactly one H2D copy later in the caller (`train.py:746`).  No device work happens here; the first device
op of the path is `common.tensor_to_vae_latent`.
"""
from __future__ import annotations

from typing import List, Optional, Union

import numpy as np
import torch

try:  # PIL is only needed for PIL inputs / outputs
    import PIL.Image
    _RESAMPLE = {"lanczos": PIL.Image.Resampling.LANCZOS, "bilinear": PIL.Image.Resampling.BILINEAR,
                 "bicubic": PIL.Image.Resampling.BICUBIC, "nearest": PIL.Image.Resampling.NEAREST}
except ImportError:  # pragma: no cover
    PIL = None
    _RESAMPLE = {}


class VaeImageProcessor:
    def __init__(self, do_resize: bool = True, vae_scale_factor: int = 8, resample: str = "lanczos",
                 do_normalize: bool = True, do_binarize: bool = False, do_convert_rgb: bool = False,
                 do_convert_grayscale: bool = False):
        if do_convert_rgb and do_convert_grayscale:
            raise ValueError("`do_convert_rgb` and `do_convert_grayscale` can not both be set to `True`")
        self.do_resize, self.vae_scale_factor, self.resample = do_resize, vae_scale_factor, resample
        self.do_normalize, self.do_binarize = do_normalize, do_binarize
        self.do_convert_rgb, self.do_convert_grayscale = do_convert_rgb, do_convert_grayscale

    # ---- elementary conversions (static in diffusers) ----
    @staticmethod
    def pil_to_numpy(images) -> np.ndarray:
        if not isinstance(images, list):
            images = [images]
        return np.stack([np.array(im).astype(np.float32) / 255.0 for im in images], axis=0)

    @staticmethod
    def numpy_to_pt(images: np.ndarray) -> torch.Tensor:
        if images.ndim == 3:
            images = images[..., None]
        return torch.from_numpy(images.transpose(0, 3, 1, 2))

    @staticmethod
    def pt_to_numpy(images: torch.Tensor) -> np.ndarray:
        return images.cpu().permute(0, 2, 3, 1).float().numpy()

    @staticmethod
    def numpy_to_pil(images: np.ndarray):
        if images.ndim == 3:
            images = images[None]
        images = (images * 255).round().astype("uint8")
        if images.shape[-1] == 1:
            return [PIL.Image.fromarray(im.squeeze(), mode="L") for im in images]
        return [PIL.Image.fromarray(im) for im in images]

    @staticmethod
    def normalize(images):
        return 2.0 * images - 1.0

    @staticmethod
    def denormalize(images):
        return (images / 2 + 0.5).clamp(0, 1)

    def get_default_height_width(self, image, height: Optional[int] = None, width: Optional[int] = None):
        if height is None:
            height = image.height if PIL is not None and isinstance(image, PIL.Image.Image) else image.shape[-2 if isinstance(image, torch.Tensor) else 1]
        if width is None:
            width = image.width if PIL is not None and isinstance(image, PIL.Image.Image) else image.shape[-1 if isinstance(image, torch.Tensor) else 2]
        return (height - height % self.vae_scale_factor, width - width % self.vae_scale_factor)

    def resize(self, image, height: Optional[int] = None, width: Optional[int] = None):
        if PIL is not None and isinstance(image, PIL.Image.Image):
            return image.resize((width, height), resample=_RESAMPLE[self.resample])
        if isinstance(image, torch.Tensor):
            return torch.nn.functional.interpolate(image, size=(height, width))
        t = torch.nn.functional.interpolate(self.numpy_to_pt(image), size=(height, width))
        return self.pt_to_numpy(t)

    def binarize(self, image):
        image[image < 0.5] = 0
        image[image >= 0.5] = 1
        return image

    def preprocess(self, image: Union["PIL.Image.Image", np.ndarray, torch.Tensor, List], height: Optional[int] = None,
                   width: Optional[int] = None) -> torch.Tensor:
        """PIL / numpy [H, W, C] in [0, 1] / torch [C, H, W] in [0, 1] (or lists / batches of them) -> float32
        [N, C, H', W'] in [-1, 1] with H', W' multiples of the VAE scale factor."""
        supported = (np.ndarray, torch.Tensor) + ((PIL.Image.Image,) if PIL is not None else ())
        if isinstance(image, supported):
            image = [image]
        elif not (isinstance(image, list) and all(isinstance(i, supported) for i in image)):
            raise ValueError(f"Input is in incorrect format: {[type(i) for i in image]}. Currently, we only support PIL image, numpy array or torch tensor")
        if PIL is not None and isinstance(image[0], PIL.Image.Image):
            if self.do_convert_rgb:
                image = [i.convert("RGB") for i in image]
            elif self.do_convert_grayscale:
                image = [i.convert("L") for i in image]
            if self.do_resize:
                height, width = self.get_default_height_width(image[0], height, width)
                image = [self.resize(i, height, width) for i in image]
            image = self.numpy_to_pt(self.pil_to_numpy(image))
        elif isinstance(image[0], np.ndarray):
            image = np.concatenate(image, axis=0) if image[0].ndim == 4 else np.stack(image, axis=0)
            image = self.numpy_to_pt(image)
            height, width = self.get_default_height_width(image, height, width)
            if self.do_resize:
                image = self.resize(image, height, width)
        else:
            image = torch.cat(image, dim=0) if image[0].ndim == 4 else torch.stack(image, dim=0)
            if self.do_convert_grayscale and image.ndim == 3:
                image = image.unsqueeze(1)
            if image.shape[1] == 4:        # latents are passed through
                return image
            height, width = self.get_default_height_width(image, height, width)
            if self.do_resize:
                image = self.resize(image, height, width)
        do_normalize = self.do_normalize
        if do_normalize and image.min() < 0:
            do_normalize = False           # already in [-1, 1] (diffusers warns and skips)
        if do_normalize:
            image = self.normalize(image)
        if self.do_binarize:
            image = self.binarize(image)
        return image

    def postprocess(self, image: torch.Tensor, output_type: str = "pil", do_denormalize: Optional[List[bool]] = None):
        if not isinstance(image, torch.Tensor):
            raise ValueError(f"Input for postprocessing is in incorrect format: {type(image)}. We only support pytorch tensor")
        if output_type == "latent":
            return image
        if do_denormalize is None:
            do_denormalize = [self.do_normalize] * image.shape[0]
        image = torch.stack([self.denormalize(image[i]) if do_denormalize[i] else image[i] for i in range(image.shape[0])])
        if output_type == "pt":
            return image
        image = self.pt_to_numpy(image)
        if output_type == "np":
            return image
        return self.numpy_to_pil(image)
