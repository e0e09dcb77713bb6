current_task = None
pending_tasks = {}


def add_task_to_queue(id_task):
    pending_tasks[id_task] = True


def start_task(id_task):
    global current_task
    current_task = id_task
    pending_tasks.pop(id_task, None)


def finish_task(id_task):
    global current_task
    if current_task == id_task:
        current_task = None
