"""Prompt-suffix vocabularies of the reference (prompts_engineering/__init__.py:1-35).  They are DATA of the file-name contract: the
sampled suffix becomes part of the prompt, the prompt becomes part of the output file name (run_aug.py:429) and so of every value in the
aug JSON -- a different list would produce different names for the same seed."""

ARTISTIC_PROMPTS = [
    "a painting of van gogh",
    "a painting of monet",
    "a painting of picasso",
    "a painting of da vinci",
    "a painting of michelangelo",
    "a painting of rembrandt",
    "a painting of raphael",
    "a painting of vermeer",
    "a painting of degas",
    "a painting of klimt",
]

IMAGE_VARIATIONS_PROMPTS = [
    "High-Speed",
    "Lens Flare",
    "HDR (High Dynamic Range)",
    "Fish-Eye Lens",
    "Black and White",
    "Long Exposure",
    "Macro",
    "Panoramic",
    "Tilt-Shift",
    "Infrared",
    "Bokeh",
    "Time-Lapse",
    "Underwater",
    "Double Exposure",
    "Sepia Tone",
    "Vintage Look",
    "Solarized",
    "Low Light",
    "Motion Blur",
    "Cross Processed",
]
