def image_grid(images, batch_size=1, rows=None):
    return images[0]


def save_image(*a, **k):
    return None, None
