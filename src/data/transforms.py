"""PIL mode helpers with the reference's names (src/data/transforms.py)."""


class RGBConvert:
    def __call__(self, img):
        return img if img.mode == "RGB" else img.convert("RGB")

    def __repr__(self):
        return type(self).__name__


class GrayscaleConvert:
    def __call__(self, img):
        return img if img.mode == "L" else img.convert("L")

    def __repr__(self):
        return type(self).__name__
