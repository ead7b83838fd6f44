"""Mode-normalising callables for torchvision-style pipelines, exported under the names the reference's loaders import
(``RGBConvert``, ``GrayscaleConvert``): every frame / guide image becomes 3-channel RGB, every mask single-channel L."""


class _ToMode:
    mode = None

    def __call__(self, image):
        if image.mode != self.mode:
            image = image.convert(self.mode)
        return image

    def __repr__(self):
        return f"{type(self).__name__}()"


RGBConvert = type("RGBConvert", (_ToMode,), {"mode": "RGB", "__doc__": "any PIL image -> RGB"})
GrayscaleConvert = type("GrayscaleConvert", (_ToMode,), {"mode": "L", "__doc__": "any PIL image -> 8-bit grey"})
