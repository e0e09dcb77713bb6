"""Host side of img2img inpainting ("whole picture" and "only masked"): mask preparation and the final overlay, as sdwui's
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
    crop: Optional[tuple] = None   # "only masked": (x1, y1, x2, y2) of the init image that is cut out, resized, repainted
    paste_to: Optional[tuple] = None   # ... and (x, y, w, h) where the result goes back into the full-size picture


def create_binary_mask(image: Image.Image, round_mask: bool = True) -> Image.Image:
    if image.mode == "RGBA" and image.getextrema()[-1] != (255, 255):
        alpha = image.split()[-1].convert("L")
        return alpha.point(lambda v: 255 if v > 128 else 0) if round_mask else alpha
    return image.convert("L")


def prepare_mask(mask: Image.Image, width: int, height: int, lat_h: int, lat_w: int, mask_blur: int = 4,
                 invert: bool = False, round_mask: bool = True) -> InpaintMask:
    m = _blurred_binary_mask(mask, mask_blur, invert, round_mask)   # blurred along x, then along y, as two separate calls
    if m.size != (width, height):
        m = m.resize((width, height), resample=Image.LANCZOS)   # resize_mode 0 ("just resize")
    overlay = Image.fromarray(np.clip(np.array(m).astype(np.float32) * 2, 0, 255).astype(np.uint8))
    lat = m.convert("RGB").resize((lat_w, lat_h))               # PIL's default resampling for RGB: bicubic
    latmask = np.moveaxis(np.array(lat, dtype=np.float32), 2, 0)[0] / 255.0
    if round_mask:
        latmask = np.around(latmask)
    return InpaintMask(torch.from_numpy(np.ascontiguousarray(latmask, dtype=np.float32)).reshape(-1), overlay, width, height, m)


def _blurred_binary_mask(mask: Image.Image, mask_blur: int, invert: bool, round_mask: bool) -> Image.Image:
    import cv2
    m = create_binary_mask(mask, round_mask)
    if invert:
        m = ImageOps.invert(m)
    if mask_blur > 0:
        k = 2 * int(2.5 * mask_blur + 0.5) + 1
        m = Image.fromarray(cv2.GaussianBlur(cv2.GaussianBlur(np.array(m), (k, 1), mask_blur), (1, k), mask_blur))
    return m


def resize_and_fill(im: Image.Image, width: int, height: int) -> Image.Image:
    """sdwui images.resize_image(resize_mode = 2): aspect-preserving LANCZOS resize centred on an RGB canvas, the bands
    left over filled by stretching the border row / column"""
    ratio, src_ratio = width / height, im.width / im.height
    src_w = width if ratio < src_ratio else im.width * height // im.height
    src_h = height if ratio >= src_ratio else im.height * width // im.width
    resized = im.resize((src_w, src_h), resample=Image.LANCZOS)
    res = Image.new("RGB", (width, height))
    res.paste(resized, box=(width // 2 - src_w // 2, height // 2 - src_h // 2))
    if ratio < src_ratio:
        band = height // 2 - src_h // 2
        if band > 0:
            res.paste(resized.resize((width, band), box=(0, 0, width, 0)), box=(0, 0))
            res.paste(resized.resize((width, band), box=(0, resized.height, width, resized.height)), box=(0, band + src_h))
    elif ratio > src_ratio:
        band = width // 2 - src_w // 2
        if band > 0:
            res.paste(resized.resize((band, height), box=(0, 0, 0, height)), box=(0, 0))
            res.paste(resized.resize((band, height), box=(resized.width, 0, resized.width, height)), box=(band + src_w, 0))
    return res


def crop_and_resize(im: Image.Image, width: int, height: int) -> Image.Image:
    """sdwui images.resize_image(resize_mode = 1)"""
    ratio, src_ratio = width / height, im.width / im.height
    src_w = width if ratio > src_ratio else im.width * height // im.height
    src_h = height if ratio <= src_ratio else im.height * width // im.width
    res = Image.new("RGB", (width, height))
    res.paste(im.resize((src_w, src_h), resample=Image.LANCZOS), box=(width // 2 - src_w // 2, height // 2 - src_h // 2))
    return res


def masked_region(mask_l: np.ndarray, pad: int):
    """sdwui masking.get_crop_region: bounding box of the non-zero mask pixels grown by `pad`, clipped to the picture;
    None when the mask is empty"""
    ys, xs = np.nonzero(mask_l)
    if ys.size == 0:
        return None
    h, w = mask_l.shape
    return (max(int(xs.min()) - pad, 0), max(int(ys.min()) - pad, 0), min(int(xs.max()) + 1 + pad, w),
            min(int(ys.max()) + 1 + pad, h))


def grow_to_aspect(region, proc_w: int, proc_h: int, img_w: int, img_h: int):
    """sdwui masking.expand_crop_region: widen or heighten the box to the processing aspect ratio, shifted back inside"""
    x1, y1, x2, y2 = region
    target = proc_w / proc_h
    if (x2 - x1) / (y2 - y1) > target:
        extra = int((x2 - x1) / target - (y2 - y1))
        y1, y2 = y1 - extra // 2, y2 + extra - extra // 2
        if y2 >= img_h:
            y1, y2 = y1 - (y2 - img_h), img_h
        if y1 < 0:
            y1, y2 = 0, y2 - y1
        y2 = min(y2, img_h)
    else:
        extra = int((y2 - y1) * target - (x2 - x1))
        x1, x2 = x1 - extra // 2, x2 + extra - extra // 2
        if x2 >= img_w:
            x1, x2 = x1 - (x2 - img_w), img_w
        if x1 < 0:
            x1, x2 = 0, x2 - x1
        x2 = min(x2, img_w)
    return x1, y1, x2, y2


def prepare_mask_only_masked(mask: Image.Image, width: int, height: int, lat_h: int, lat_w: int, mask_blur: int = 4,
                             invert: bool = False, padding: int = 32, round_mask: bool = True) -> Optional[InpaintMask]:
    """inpaint_full_res ("Only masked", sdwui StableDiffusionProcessingImg2Img.init): the padded bounding box of the mask,
    grown to the processing aspect ratio, is cut out of the FULL-SIZE init image and repainted at (width, height); the
    overlay keeps the full-size picture.  None when the mask is empty (sdwui then runs plain img2img)."""
    m = _blurred_binary_mask(mask, mask_blur, invert, round_mask).convert("L")
    region = masked_region(np.array(m), padding)
    if region is None:
        return None
    x1, y1, x2, y2 = grow_to_aspect(region, width, height, m.width, m.height)
    work = resize_and_fill(m.crop((x1, y1, x2, y2)), width, height)        # RGB, processing size
    lat = np.moveaxis(np.array(work.resize((lat_w, lat_h)), dtype=np.float32), 2, 0)[0] / 255.0
    if round_mask:
        lat = np.around(lat)
    return InpaintMask(torch.from_numpy(np.ascontiguousarray(lat, dtype=np.float32)).reshape(-1), m, width, height,
                       work.convert("L"), (x1, y1, x2, y2), (x1, y1, x2 - x1, y2 - y1))


def crop_init_images(images: List[Image.Image], mask: InpaintMask) -> torch.Tensor:
    """"only masked": every full-size init image -> its crop region at the processing size, uint8 [b, H, W, 3]"""
    out = [np.array(resize_and_fill(im.convert("RGB").crop(mask.crop), mask.width, mask.height)) for im in images]
    return torch.from_numpy(np.stack(out))


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


def apply_overlays(images_u8: torch.Tensor, overlays: Optional[List[Image.Image]], paste_to=None) -> torch.Tensor:
    """generated uint8 [b, H, W, 3] (host) + overlays -> uint8 [b, H', W', 3]: sdwui apply_overlay —
    image.convert('RGBA').alpha_composite(overlay); with `paste_to` = (x, y, w, h) ("only masked") the generated picture is
    first resized to (w, h) and pasted at (x, y) of an empty canvas of the overlay's (= the full init image's) size."""
    if not overlays:
        return images_u8
    out = []
    for k in range(images_u8.shape[0]):
        img = Image.fromarray(images_u8[k].numpy(), "RGB")
        overlay = overlays[k % len(overlays)]
        if paste_to is not None:
            x, y, w, h = paste_to
            base = Image.new("RGBA", (overlay.width, overlay.height))
            base.paste(crop_and_resize(img, w, h), (x, y))
            img = base
        img = img.convert("RGBA")
        img.alpha_composite(overlay)
        out.append(torch.from_numpy(np.array(img.convert("RGB"))))
    return torch.stack(out)
