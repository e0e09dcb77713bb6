"""Host side of img2img inpainting ("whole picture", fill = original): mask preparation and the final overlay, as sdwui's
StableDiffusionProcessingImg2Img.init / apply_overlay do them (PIL + OpenCV on the host; none of it is per-step work).

Reference boundary: the extension forwards `image_mask` as the API field `mask` together with `mask_blur`,
`inpainting_fill`, `inpaint_full_res`, `inpainting_mask_invert` (scripts/spartan/worker.py:365-373, :406-410); the remote
sdwui then runs this logic.  Upstream symbols followed: modules/masking.py / processing.py `create_binary_mask`, the
per-axis `cv2.GaussianBlur`, the x2-clipped overlay mask, the rounded latent mask, `apply_overlay`.
"""
from dataclasses import dataclass
from typing import List, Optional

import numpy as np
import torch
from PIL import Image, ImageFilter, ImageOps


@dataclass
class InpaintMask:
    latmask: torch.Tensor          # fp32 [lat_h * lat_w]: 1 = repaint, 0 = keep the original latent
    overlay_mask: Image.Image      # 'L', image resolution: where the ORIGINAL pixels are pasted back (inverted alpha)
    width: int
    height: int
    fill_mask: Optional[Image.Image] = None   # 'L', image resolution: the blurred mask itself (inpainting_fill 0)


def create_binary_mask(image: Image.Image, round_mask: bool = True) -> Image.Image:
    if image.mode == "RGBA" and image.getextrema()[-1] != (255, 255):
        alpha = image.split()[-1].convert("L")
        return alpha.point(lambda v: 255 if v > 128 else 0) if round_mask else alpha
    return image.convert("L")


def prepare_mask(mask: Image.Image, width: int, height: int, lat_h: int, lat_w: int, mask_blur: int = 4,
                 invert: bool = False, round_mask: bool = True) -> InpaintMask:
    import cv2
    m = create_binary_mask(mask, round_mask)
    if invert:
        m = ImageOps.invert(m)
    if mask_blur > 0:   # blurred along x, then along y, as two separate calls
        arr = np.array(m)
        k = 2 * int(2.5 * mask_blur + 0.5) + 1
        arr = cv2.GaussianBlur(arr, (k, 1), mask_blur)
        arr = cv2.GaussianBlur(arr, (1, k), mask_blur)
        m = Image.fromarray(arr)
    if m.size != (width, height):
        m = m.resize((width, height), resample=Image.LANCZOS)   # resize_mode 0 ("just resize")
    overlay = Image.fromarray(np.clip(np.array(m).astype(np.float32) * 2, 0, 255).astype(np.uint8))
    lat = m.convert("RGB").resize((lat_w, lat_h))               # PIL's default resampling for RGB: bicubic
    latmask = np.moveaxis(np.array(lat, dtype=np.float32), 2, 0)[0] / 255.0
    if round_mask:
        latmask = np.around(latmask)
    return InpaintMask(torch.from_numpy(np.ascontiguousarray(latmask, dtype=np.float32)).reshape(-1), overlay, width, height, m)


def fill_masked(init_images_u8: torch.Tensor, mask: InpaintMask) -> torch.Tensor:
    """inpainting_fill = 0 ("fill", sdwui modules/masking.py fill): the masked region of every init image is replaced by a
    cascade of blurs of its surroundings before the image is encoded.  uint8 [b, H, W, 3] -> same."""
    m = mask.fill_mask
    out = torch.empty_like(init_images_u8)
    for k in range(init_images_u8.shape[0]):
        image = Image.fromarray(init_images_u8[k].cpu().numpy(), "RGB")
        image_mod = Image.new("RGBA", (image.width, image.height))
        image_masked = Image.new("RGBa", (image.width, image.height))
        image_masked.paste(image.convert("RGBA").convert("RGBa"), mask=ImageOps.invert(m.convert("L")))
        image_masked = image_masked.convert("RGBa")
        for radius, repeats in [(256, 1), (64, 1), (16, 2), (4, 4), (2, 2), (0, 1)]:
            blurred = image_masked.filter(ImageFilter.GaussianBlur(radius)).convert("RGBA")
            for _ in range(repeats):
                image_mod.alpha_composite(blurred)
        out[k] = torch.from_numpy(np.array(image_mod.convert("RGB")))
    return out


def overlays_for(init_images_u8: torch.Tensor, mask: InpaintMask) -> List[Image.Image]:
    """per init image: the original pixels with alpha = 255 - overlay_mask (premultiplied paste, then back to RGBA)"""
    out = []
    inv = ImageOps.invert(mask.overlay_mask.convert("L"))
    for k in range(init_images_u8.shape[0]):
        image = Image.fromarray(init_images_u8[k].cpu().numpy(), "RGB")
        masked = Image.new("RGBa", (image.width, image.height))
        masked.paste(image.convert("RGBA").convert("RGBa"), mask=inv)
        out.append(masked.convert("RGBA"))
    return out


def apply_overlays(images_u8: torch.Tensor, overlays: Optional[List[Image.Image]]) -> torch.Tensor:
    """generated uint8 [b, H, W, 3] (host) + overlays -> uint8 [b, H, W, 3]: image.convert('RGBA').alpha_composite(overlay)"""
    if not overlays:
        return images_u8
    out = torch.empty_like(images_u8)
    for k in range(images_u8.shape[0]):
        img = Image.fromarray(images_u8[k].numpy(), "RGB").convert("RGBA")
        img.alpha_composite(overlays[k % len(overlays)])
        out[k] = torch.from_numpy(np.array(img.convert("RGB")))
    return out
