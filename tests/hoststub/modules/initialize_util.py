def gradio_server_name():
    return None
